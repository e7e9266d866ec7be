"""BASELINE.json config 5: the fps_benchmark_demo.py protocol (3 rounds x n_iter, forward only under no_grad, CUDA
events; reference: fps_benchmark_demo.py:53-66) over 50k/100k/300k/1M splats x 720p/1080p/4K, with and without the
per-frame mesh-frame update inside the timed region.  One JSON line per cell -> profiles/<round>/fps_sweep.jsonl."""
import json, os, sys, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
from gaussianavatars_b200 import synthetic as syn, rasterizer as R
from gaussianavatars_b200.model import MeshBoundGaussians
from gaussianavatars_b200.renderer import render

dev = torch.device("cuda:0")
class Pipe: debug=False; compute_cov3D_python=False; convert_SHs_python=False
peak = 6585.8
try: peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception: pass
sizes = [int(x) for x in os.environ.get("SWEEP_P", "50000,100000,300000,1000000").split(",")]
res = {"720p": (1280, 720), "1080p": (1920, 1080), "4K": (3840, 2160)}
n_iter = int(os.environ.get("SWEEP_ITERS", "100"))
verts, faces = syn.head_mesh()
bg = torch.ones(3, device=dev)
R.keep_last_state(True)
for P in sizes:
    params = syn.avatar_splats(P, n_faces=faces.shape[0], seed=0, sh_degree=3)
    pc = MeshBoundGaussians(params, 3, verts, faces, device=dev)
    pc.select_mesh_by_timestep(0)
    for name, (W, H) in res.items():
        cam = syn.orbit_camera(W, H).to(dev)
        for with_mesh in (False, True):
            fps = []
            with torch.no_grad():
                for _ in range(5): render(cam, pc, Pipe, bg)
                _, _, _, n = R.export_last_binning()
                for rnd in range(3):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(n_iter):
                        if with_mesh: pc.update_mesh_properties(pc.verts)
                        render(cam, pc, Pipe, bg)["render"]
                    e1.record(); torch.cuda.synchronize()
                    fps.append(n_iter / (e0.elapsed_time(e1) / 1e3))
            fps.sort()
            alg = P * 240 + P * 48 + n * 12 + n * 24 + n * 40 + H * W * 12  # SURVEY 8(d) forward (inference) bytes
            gbs = alg * fps[1] / 1e9
            print(json.dumps({"splats": P, "res": name, "W": W, "H": H, "instances": int(n), "mesh_update_in_loop": with_mesh,
                              "fps_median": round(fps[1], 1), "fps_best": round(fps[2], 1), "ms_median": round(1e3 / fps[1], 4),
                              "algorithmic_GBps": round(gbs, 1), "frac_of_measured_hbm_peak": round(gbs / peak, 4)}), flush=True)
    del pc
    torch.cuda.empty_cache()
