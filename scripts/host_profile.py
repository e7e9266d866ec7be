"""Where the HOST time of a training step goes (config 3 at the reference's 550x802: the GPU work is short enough there
that the Python/launch path is the limit).  Section timers (no syncs added) + cProfile."""
import cProfile, io, json, os, pstats, sys, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
import gaussianavatars_b200 as g
from gaussianavatars_b200 import synthetic as syn, _native as N
from gaussianavatars_b200.model import MeshBoundGaussians
from gaussianavatars_b200.renderer import render

dev = torch.device("cuda:0")
class Pipe: debug=False; compute_cov3D_python=False; convert_SHs_python=False
P, W, H = int(os.environ.get("P", 150000)), int(os.environ.get("W", 550)), int(os.environ.get("H", 802))
verts, faces = syn.head_mesh()
params = syn.avatar_splats(P, n_faces=faces.shape[0], seed=0, sh_degree=3)
bg = torch.ones(3, device=dev)
cams = [syn.orbit_camera(W, H, azimuth_deg=-60 + 120 * (i + .5) / 16).to(dev) for i in range(16)]
gts = [torch.randint(0, 256, (3, H, W), dtype=torch.uint8, device=dev) for _ in range(2)]
pc = MeshBoundGaussians(params, 3, verts, faces, pose_fn=syn.pose_mesh, device=dev, requires_grad=True)
opt = g.Adam([{"params": [p], "lr": 1e-4, "name": str(i)} for i, p in enumerate(pc.parameters())], lr=0.0, eps=1e-15)
posed = [syn.pose_mesh(pc.verts_rest, i).contiguous().requires_grad_(True) for i in range(16)]
sec = {k: 0.0 for k in ("zero_grad", "mesh", "render", "loss", "backward", "adam")}
def step(i, timed=True):
    t = [time.perf_counter()]
    opt.zero_grad(set_to_none=True); t.append(time.perf_counter())
    pc.update_mesh_properties(posed[i % 16]); t.append(time.perf_counter())
    out = render(cams[i % 16], pc, Pipe, bg); t.append(time.perf_counter())
    loss = g.photometric_loss(out["render"], gts[i % 2], 0.2); t.append(time.perf_counter())
    loss.backward(); t.append(time.perf_counter())
    opt.step(); t.append(time.perf_counter())
    if timed:
        for k, a, b in zip(sec, t[:-1], t[1:]): sec[k] += b - a
for i in range(20): step(i, False)
torch.cuda.synchronize()
K = 300
N.host_times(True)
t0 = time.perf_counter()
for i in range(K): step(i)
t_enq = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
ht = N.host_times(True)
print(json.dumps({"W": W, "H": H, "P": P, "us_per_step_total": round(t_all / K * 1e6, 1), "us_per_step_host_enqueue": round(t_enq / K * 1e6, 1),
                  "host_us_by_section": {k: round(v / K * 1e6, 1) for k, v in sec.items()},
                  "inside_gab200_forward_us": ht}), flush=True)
# per-stage GPU time of the same step (library events on: perturbs the step a little) and which depth-sort path ran
from gaussianavatars_b200 import rasterizer as R
N.stage_timing(True); N.stage_times(True)
for i in range(48): step(i, False)
st_ = N.stage_times(True); N.stage_timing(False)
R.keep_last_state(True); step(0, False); torch.cuda.synchronize()
print(json.dumps({"gpu_stage_us": {k: round(v[0] / max(v[1], 1) * 1e3, 1) for k, v in st_.items()},
                  "gpu_stage_sum_us": round(sum(v[0] / max(v[1], 1) for v in st_.values()) * 1e3, 1),
                  "depth_sort_path": int(R._last[1].depth_sort_path), "N": int(R._last[1].num_rendered),
                  "depth_sort_env": os.environ.get("GAB200_DEPTH_SORT", "default")}), flush=True)
R.keep_last_state(False)
if os.environ.get("NOPROF") == "1":
    sys.exit(0)
pr = cProfile.Profile(); pr.enable()
for i in range(K): step(i, False)
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45); print(s.getvalue()[:9000])
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(30); print(s.getvalue()[:6000])
