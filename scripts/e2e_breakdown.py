"""Where the e2e step's extra time over the resident step goes: replay time of GraphedFrame variants (events around
K back-to-back replays, no L2 flush): resident dL_dimage | + l1 loss | + host camera block + loss read-back |
+ ground-truth upload prefetched on the copy stream."""
import json, os, sys, math
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
from gaussianavatars_b200 import synthetic as syn
from gaussianavatars_b200.model import MeshBoundGaussians
from gaussianavatars_b200.graph import GraphedFrame, camera_block

dev = torch.device("cuda:0")
P, W, H = 100000, 1920, 1080
verts, faces = syn.head_mesh()
params = syn.avatar_splats(P, n_faces=faces.shape[0], seed=0, sh_degree=3)
cams = [syn.orbit_camera(W, H, r=1.0, fovy_deg=20.0, azimuth_deg=-60 + 120 * (i + .5) / 16, elevation_deg=5 * math.sin(i)) for i in range(16)]
blocks = [camera_block(c) for c in cams]
bg = torch.ones(3, device=dev)
gout = torch.randn(3, H, W, generator=torch.Generator().manual_seed(1)).to(dev) / (3 * H * W)
gt = torch.randint(0, 256, (3, H, W), dtype=torch.uint8)
gt_pin = gt.pin_memory()
K = 60

def timed(fr, per_step=None):
    for i in range(5):
        if per_step: per_step(i)
        fr.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        if per_step: per_step(i)
        fr.run()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K

res = {}
for name, kw in [("resident_dL_dimage", dict(loss="dL_dimage")), ("resident_l1_u8", dict(loss="l1_u8")),
                 ("host_cam_l1_u8_gt_resident", dict(loss="l1_u8", host_inputs=True)),
                 ("host_cam_l1_u8_gt_uploaded", dict(loss="l1_u8", host_inputs=True))]:
    pc = MeshBoundGaussians(params, 3, verts, faces, pose_fn=syn.pose_mesh, device=dev, requires_grad=True)
    host = kw.get("host_inputs", False)
    fr = GraphedFrame(pc, W, H, cams[0].FoVx, cams[0].FoVy, bg, warm_cameras=[b.pin_memory() if host else b.to(dev) for b in blocks], **kw)
    fr.set_inputs(camera=blocks[0], verts=pc.verts_rest, gt_u8=None if kw["loss"] == "dL_dimage" else gt.to(dev),
                  dL_dimage=gout if kw["loss"] == "dL_dimage" else None)
    fr.capture()
    step = None
    if name.endswith("uploaded"):
        step = lambda i: fr.set_inputs(gt_u8=gt_pin)
    res[name] = round(timed(fr, step), 4)
    assert not fr.overflowed()
    del fr, pc
print(json.dumps(res))
