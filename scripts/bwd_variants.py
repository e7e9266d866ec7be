"""Backward-blend schedule sweep on the headline workload (100k splats, 1080p): per variant of
GAB200_TUNE_BWD_VARIANT x heavy-tile threshold, the blend_bwd stage time (library stage events), the whole eager step,
and the largest gradient difference against variant 0 (same arithmetic, different summation order)."""
import json, os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
from gaussianavatars_b200 import synthetic as syn, _native as N, rasterizer as R
from gaussianavatars_b200.model import MeshBoundGaussians
from gaussianavatars_b200.renderer import render

dev = torch.device("cuda:0")
class Pipe: debug = False; compute_cov3D_python = False; convert_SHs_python = False
P, W, H = int(os.environ.get("P", 100000)), int(os.environ.get("W", 1920)), int(os.environ.get("H", 1080))
verts, faces = syn.head_mesh()
params = syn.avatar_splats(P, n_faces=faces.shape[0], seed=0, sh_degree=3)
pc = MeshBoundGaussians(params, 3, verts, faces, pose_fn=syn.pose_mesh, device=dev, requires_grad=True)
import math
cams = [syn.orbit_camera(W, H, r=1.0, fovy_deg=20.0, azimuth_deg=-60 + 120 * (i + .5) / 16, elevation_deg=5 * math.sin(i)).to(dev) for i in range(16)]
posed = [syn.pose_mesh(pc.verts_rest, i).contiguous().requires_grad_(True) for i in range(16)]
bg = torch.ones(3, device=dev)
gout = torch.randn(3, H, W, generator=torch.Generator().manual_seed(1)).to(dev) / (3 * H * W)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def step(i):
    for p in pc.parameters(): p.grad = None
    pc.update_mesh_properties(posed[i % 16])
    out = render(cams[i % 16], pc, Pipe, bg)
    out["render"].backward(gout)

ref = None
variants = [int(x) for x in os.environ.get("VARIANTS", "0,1,2,3").split(",")]
heavies = [int(x) for x in os.environ.get("HEAVY", "512,1024,2048").split(",")]
for v in variants:
    for hv in heavies:
        N.tune(N.TUNE_BWD_VARIANT, v); N.tune(N.TUNE_HEAVY_BWD, hv)
        for i in range(6): step(i)
        torch.cuda.synchronize()
        step(0); torch.cuda.synchronize()
        g = torch.cat([p.grad.reshape(-1) for p in pc.parameters()]).clone()
        if ref is None: ref = g
        diff = float((g - ref).abs().max() / ref.abs().max())
        N.stage_timing(True); N.stage_times(True)
        K = 40
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tot = 0.0
        for i in range(K):
            flush.fill_(i & 255)
            e0.record(); step(i); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
        st = N.stage_times(True); N.stage_timing(False)
        print(json.dumps({"variant": v, "heavy_bwd": hv, "blend_bwd_us": round(st["blend_bwd"][0] / st["blend_bwd"][1] * 1e3, 1),
                          "blend_fwd_us": round(st["blend_fwd"][0] / st["blend_fwd"][1] * 1e3, 1),
                          "step_ms_eager_with_events": round(tot / K, 4), "max_grad_diff_vs_first": diff}), flush=True)
