"""BASELINE.json config 3: bound avatar, 150k splats, 16 cameras, one --bind_to_mesh training step per camera:
face frame -> fused forward -> (1-l) L1 + l (1-SSIM) -> backward (incl. dL/dverts) -> Adam on the six splat arrays.

Two arms per resolution, same rasterizer underneath:
  ours   : gaussianavatars_b200.photometric_loss (uint8 ground truth, 2 launches) + gaussianavatars_b200.Adam (1 launch)
  eager  : the reference's loss written with torch ops (5 grouped conv2d + autograd, float32 ground truth) and
           torch.optim.Adam(eps=1e-15) as scene/gaussian_model.py:222 builds it
and the two new kernels timed alone against the measured HBM peak.  One JSON line per measurement."""
import json, math, os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
import torch.nn.functional as F
import gaussianavatars_b200 as g
from gaussianavatars_b200 import synthetic as syn
from gaussianavatars_b200.model import MeshBoundGaussians
from gaussianavatars_b200.renderer import render

dev = torch.device("cuda:0")
LAMBDA = 0.2
class Pipe: debug=False; compute_cov3D_python=False; convert_SHs_python=False
P = int(os.environ.get("P", 150000))
PEAK = 6585.8
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass

_w1 = torch.tensor([math.exp(-(x - 5) ** 2 / (2 * 1.5 ** 2)) for x in range(11)])
_w1 = _w1 / _w1.sum()
WINDOW = (_w1[:, None] @ _w1[None, :]).expand(3, 1, 11, 11).contiguous().to(dev)

def eager_loss(img, gt):
    conv = lambda a: F.conv2d(a, WINDOW, padding=5, groups=3)
    mu1, mu2 = conv(img), conv(gt)
    mu1_sq, mu2_sq, mu12 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    s1, s2, s12 = conv(img * img) - mu1_sq, conv(gt * gt) - mu2_sq, conv(img * gt) - mu12
    smap = ((2 * mu12 + 1e-4) * (2 * s12 + 9e-4)) / ((mu1_sq + mu2_sq + 1e-4) * (s1 + s2 + 9e-4))
    return torch.abs(img - gt).mean() * (1 - LAMBDA) + (1 - smap.mean()) * LAMBDA

def timed(fn, iters, warm):
    for i in range(warm): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

verts, faces = syn.head_mesh()
params = syn.avatar_splats(P, n_faces=faces.shape[0], seed=0, sh_degree=3)
bg = torch.ones(3, device=dev)
for (W, H) in ((550, 802), (1920, 1080)):
    cams = [syn.orbit_camera(W, H, azimuth_deg=-60 + 120 * (i + .5) / 16).to(dev) for i in range(16)]
    gts_u8 = [torch.randint(0, 256, (3, H, W), dtype=torch.uint8, device=dev) for _ in range(2)]
    gts_f = [t.float() / 255 for t in gts_u8]
    res = {}
    for arm in (("ours",) if os.environ.get("ONLY_OURS") == "1" else ("ours", "eager")):
        pc = MeshBoundGaussians(params, 3, verts, faces, pose_fn=syn.pose_mesh, device=dev, requires_grad=True)
        groups = [{"params": [p], "lr": 1e-4, "name": str(i)} for i, p in enumerate(pc.parameters())]
        opt = (g.Adam if arm == "ours" else torch.optim.Adam)(groups, lr=0.0, eps=1e-15)
        posed = [syn.pose_mesh(pc.verts_rest, i).contiguous().requires_grad_(True) for i in range(16)]
        def step(i):
            opt.zero_grad(set_to_none=True)
            pc.update_mesh_properties(posed[i % 16])
            out = render(cams[i % 16], pc, Pipe, bg)
            if arm == "ours":
                loss = g.photometric_loss(out["render"], gts_u8[i % 2], LAMBDA)
            else:
                loss = eager_loss(out["render"][None], gts_f[i % 2][None])
            loss.backward()          # reaches the vertices through the face-frame kernel's backward
            opt.step()
        res[arm] = timed(step, 64, 8)
    print(json.dumps({"config": "BASELINE configs[2]: 150k bound splats, 16 cameras, training step (L1 + 0.2 D-SSIM, Adam)",
                      "splats": P, "W": W, "H": H, "ms_per_training_step": round(res["ours"], 4),
                      "steps_per_s": round(1e3 / res["ours"], 1),
                      "ms_per_step_with_eager_loss_and_torch_adam": round(res["eager"], 4) if "eager" in res else None}), flush=True)
    if os.environ.get("ONLY_OURS") == "1":
        continue

    # ---- the loss alone (forward + gradient), ours vs eager
    img = torch.rand(3, H, W, device=dev)
    def ours_loss(i):
        x = img.requires_grad_(True); x.grad = None
        g.photometric_loss(x, gts_u8[i % 2], LAMBDA).backward()
    def eager(i):
        x = img.requires_grad_(True); x.grad = None
        eager_loss(x[None], gts_f[i % 2][None]).backward()
    t_ours, t_eager = timed(ours_loss, 50, 5), timed(eager, 20, 3)
    n = 3 * H * W
    algo = n * (4 + 1 + 12) + n * (12 + 4 + 1 + 4)      # stats: img+gt in, 3 maps out; grad: 3 maps + img + gt in, grad out
    print(json.dumps({"kernel": "photometric_loss fwd+grad (2 launches + autograd glue)", "W": W, "H": H, "us": round(t_ours * 1e3, 1),
                      "eager_torch_us": round(t_eager * 1e3, 1), "algorithmic_bytes": algo,
                      "achieved_gbs": round(algo / t_ours / 1e6, 1), "frac_of_hbm_peak": round(algo / t_ours / 1e6 / PEAK, 4)}), flush=True)

if os.environ.get("ONLY_OURS") == "1":
    sys.exit(0)
# ---- Adam alone at 150k splats x 59 floats (flat views like the fused backward's gradient buffer)
sizes = [P * 3, P * 3, P * 45, P, P * 3, P * 4]
for name, cls in (("ours", g.Adam), ("torch_default", torch.optim.Adam), ("torch_fused", lambda gr, **k: torch.optim.Adam(gr, fused=True, **k))):
    ps = [torch.nn.Parameter(torch.randn(s, device=dev)) for s in sizes]
    opt = cls([{"params": [p], "lr": 1e-3} for p in ps], lr=0.0, eps=1e-15)
    for p in ps: p.grad = torch.randn_like(p)
    t = timed(lambda i: opt.step(), 50, 5)
    nbytes = sum(sizes) * 28
    print(json.dumps({"kernel": f"adam step, {name}", "elements": sum(sizes), "us": round(t * 1e3, 1), "algorithmic_bytes": nbytes,
                      "achieved_gbs": round(nbytes / t / 1e6, 1), "frac_of_hbm_peak": round(nbytes / t / 1e6 / PEAK, 4)}), flush=True)
