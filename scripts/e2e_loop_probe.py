"""e2e loop of bench.py taken apart: GPU time per step (events) and HOST time per step (perf_counter) for
  full        two alternating GraphedFrames, uploads of the next step on the copy stream, loss read one step late
  no_readback the same without waiting for / reading the previous step's loss
  no_upload   the same without the per-step H2D uploads (inputs stay resident)
  one_frame   a single frame, no uploads (the resident l1 graph back to back)."""
import json, math, os, sys, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
from gaussianavatars_b200 import synthetic as syn
from gaussianavatars_b200.model import MeshBoundGaussians
from gaussianavatars_b200.graph import GraphedFrame, camera_block

dev = torch.device("cuda:0")
P, W, H, K = 100000, 1920, 1080, 60
verts, faces = syn.head_mesh()
params = syn.avatar_splats(P, n_faces=faces.shape[0], seed=0, sh_degree=3)
cams = [syn.orbit_camera(W, H, r=1.0, fovy_deg=20.0, azimuth_deg=-60 + 120 * (i + .5) / 16, elevation_deg=5 * math.sin(i)) for i in range(16)]
blocks = [camera_block(c).pin_memory() for c in cams]
bg = torch.ones(3, device=dev)
gt_pin = [torch.randint(0, 256, (3, H, W), dtype=torch.uint8).pin_memory() for _ in range(2)]
pc = MeshBoundGaussians(params, 3, verts, faces, pose_fn=syn.pose_mesh, device=dev, requires_grad=True)
posed = [syn.pose_mesh(pc.verts_rest, i).contiguous() for i in range(16)]
frames = []
for k in range(2):
    f = GraphedFrame(pc, W, H, cams[0].FoVx, cams[0].FoVy, bg, loss="l1_u8", host_inputs=True, warm_cameras=blocks)
    f.set_inputs(camera=blocks[0], verts=posed[0], gt_u8=gt_pin[k])
    f.capture()
    frames.append(f)
done = [torch.cuda.Event() for _ in range(2)]

def loop(upload, readback, two=True):
    host = 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(2):   # first repetition = warm-up
        torch.cuda.synchronize()
        host = 0.0
        e0.record()
        for i in range(K):
            t0 = time.perf_counter()
            f = frames[i % 2] if two else frames[0]
            f.set_inputs(verts=posed[i % 16])
            f.run()
            done[i % 2].record()
            if upload:
                g = frames[(i + 1) % 2] if two else frames[0]
                g.set_inputs(camera=blocks[(i + 1) % 16], gt_u8=gt_pin[(i + 1) % 2])
            t1 = time.perf_counter()
            host += t1 - t0
            if readback and i > 0:
                done[(i - 1) % 2].synchronize()
                float(frames[(i - 1) % 2].loss_host)
        e1.record()
        torch.cuda.synchronize()
    return {"gpu_ms_per_step": round(e0.elapsed_time(e1) / K, 4), "host_enqueue_us_per_step": round(host / K * 1e6, 1)}

out = {"full": loop(True, True), "no_readback": loop(True, False), "no_upload": loop(False, True),
       "one_frame_no_upload": loop(False, False, two=False)}
print(json.dumps(out))
