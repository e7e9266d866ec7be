"""Scratch timing of the stages at the headline workload (not the bench contract; see bench.py)."""
import os, sys, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
from gaussianavatars_b200 import synthetic as syn, rasterizer as R
from gaussianavatars_b200.model import MeshBoundGaussians
from gaussianavatars_b200.renderer import render

dev = torch.device("cuda:0")
W, H, P = int(os.environ.get("W", 1920)), int(os.environ.get("H", 1080)), int(os.environ.get("P", 100000))
verts, faces = syn.head_mesh()
params = syn.avatar_splats(P, n_faces=faces.shape[0], seed=0, sh_degree=3, scale_gain=float(os.environ.get("GAIN", 1.0)))
cam = syn.orbit_camera(W, H)
class Pipe: debug=False; compute_cov3D_python=False; convert_SHs_python=False
bg = torch.ones(3, device=dev)
gout = torch.randn(3, H, W, device=dev)
FAST = os.environ.get('FAST') == '1'
for exact in ((False,) if FAST else (True, False)):
    R.set_exact_binning(exact); R.keep_last_state(True)
    pc = MeshBoundGaussians(params, 3, verts, faces, device=dev, requires_grad=True)
    pc.update_mesh_properties(pc.verts_rest.clone().requires_grad_(True))
    for fused in ((True,) if FAST else (True, False)):
        def step(bw=True):
            out = render(cam, pc, Pipe, bg, fused=fused)
            if bw:
                for p_ in pc.parameters(): p_.grad = None
                out["render"].backward(gout)
            return out
        for _ in range(5): out = step()
        torch.cuda.synchronize()
        _, _, _, n = R.export_last_binning()
        for bw in (False, True):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            t0 = time.perf_counter(); ev[0].record()
            iters = 30
            for _ in range(iters):
                if bw: step(True)
                else:
                    with torch.no_grad(): step(False)
            ev[1].record(); torch.cuda.synchronize(); t1 = time.perf_counter()
            if bw:
                from gaussianavatars_b200 import _native as NN
                NN.stage_timing(True); NN.stage_times(True)
                for _ in range(10): step(True)
                st_ = NN.stage_times(True); NN.stage_timing(False)
                print("   stages us:", {k: round(v[0] / max(v[1], 1) * 1e3, 1) for k, v in st_.items()}, flush=True)
            print(f"exact={exact} fused={fused} bw={bw} N={n} vis={int((out['radii']>0).sum())} "
                  f"gpu {ev[0].elapsed_time(ev[1])/iters*1e3:.1f} us/frame wall {(t1-t0)/iters*1e6:.1f} us/frame", flush=True)
