"""BASELINE.json config 3 (bound avatar, 150k splats, 16 cameras, --bind_to_mesh training step) with everything this
round added: per step  pose -> [one CUDA graph: face frame, fused forward, (1-l) L1 + l (1-SSIM) + position/scale
regularisers, backward down to the vertices] -> Adam on the six splat arrays (one launch).  Compared with the same step
run eagerly through render() + autograd.  One JSON line per resolution."""
import json, math, os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
import gaussianavatars_b200 as g
from gaussianavatars_b200 import synthetic as syn
from gaussianavatars_b200.graph import GraphedFrame, camera_block
from gaussianavatars_b200.model import MeshBoundGaussians
from gaussianavatars_b200.renderer import render

dev = torch.device("cuda:0")
class Pipe: debug = False; compute_cov3D_python = False; convert_SHs_python = False
P, K = int(os.environ.get("P", 150000)), int(os.environ.get("ITERS", 64))
verts, faces = syn.head_mesh()
params = syn.avatar_splats(P, n_faces=faces.shape[0], seed=0, sh_degree=3)
bg = torch.ones(3, device=dev)

def timed(fn):
    for i in range(8): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K

for (W, H) in ((550, 802), (1920, 1080)):
    cams = [syn.orbit_camera(W, H, azimuth_deg=-60 + 120 * (i + .5) / 16, elevation_deg=5 * math.sin(i)) for i in range(16)]
    gts = [torch.randint(0, 256, (3, H, W), dtype=torch.uint8, device=dev) for _ in range(2)]
    res = {"config": "3", "splats": P, "W": W, "H": H}
    for arm in ("eager", "graph"):
        pc = MeshBoundGaussians(params, 3, verts, faces, pose_fn=syn.pose_mesh, device=dev, requires_grad=True)
        opt = g.Adam([{"params": [p], "lr": 1e-4, "name": str(i)} for i, p in enumerate(pc.parameters())], lr=0.0, eps=1e-15)
        posed = [syn.pose_mesh(pc.verts_rest, i).contiguous() for i in range(16)]
        if arm == "eager":
            cd = [c.to(dev) for c in cams]
            def step(i):
                opt.zero_grad(set_to_none=True)
                v = posed[i % 16].requires_grad_(True)
                pc.update_mesh_properties(v)
                out = render(cd[i % 16], pc, Pipe, bg)
                loss = g.photometric_loss(out["render"], gts[i % 2], 0.2)
                lx, ls = g.binding_regularizers(pc._xyz, pc._scaling, out["radii"], pc.binding, pc.face_scaling)
                (loss + lx + ls).backward()
                opt.step()
        else:
            blocks = [camera_block(c).to(dev) for c in cams]
            fr = GraphedFrame(pc, W, H, cams[0].FoVx, cams[0].FoVy, bg, loss="photometric", lambda_dssim=0.2, regularizers={},
                              warm_cameras=blocks)
            fr.set_inputs(camera=blocks[0], verts=posed[0], gt_u8=gts[0])
            fr.capture()
            def step(i):
                fr.set_inputs(camera=blocks[i % 16], verts=posed[i % 16], gt_u8=gts[i % 2])
                fr.run()
                opt.step()
        res[arm + "_ms_per_step"] = round(timed(step), 4)
        if arm == "graph":
            res["graph_overflow"] = fr.overflowed()
            res["loss_finite"] = bool(torch.isfinite(fr.loss))
    print(json.dumps(res), flush=True)
