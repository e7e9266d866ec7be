"""Why the CPU arm's binding getters went from 19 ms to 3 s: time them and the C oracle under different thread
settings on the GPU box's host (no GPU used)."""
import json, os, sys, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
import numpy as np
sys.argv = ["bench.py"]
import bench
from gaussianavatars_b200 import synthetic as syn
from oracle import rasterizer as orc

info = {"cpu_count": os.cpu_count(), "affinity": len(os.sched_getaffinity(0)), "torch_threads": torch.get_num_threads(),
        "env": {k: v for k, v in os.environ.items() if "OMP" in k or "MKL" in k or "KMP" in k}}
try:
    info["cpu.max"] = open("/sys/fs/cgroup/cpu.max").read().strip()
except Exception as e:
    info["cpu.max"] = repr(e)
print(json.dumps(info), flush=True)
verts, faces = syn.head_mesh()
params = syn.avatar_splats(100000, n_faces=faces.shape[0], seed=0, sh_degree=3)
cams = bench.make_cameras(16)

def run(tag, frames=3):
    t, tb = bench.cpu_frames(params, verts, faces, cams, frames, threads=None)
    print(json.dumps({"tag": tag, "frame_ms": [round(x * 1e3, 1) for x in t], "binding_ms": [round(x * 1e3, 1) for x in tb],
                      "torch_threads": torch.get_num_threads()}), flush=True)

# 1. exactly what bench.py does now
run("bench default (orc.set_threads(affinity))")
for nt in (64, 32, 16):
    orc.set_threads(nt)
    _orig = bench.host_cpus
    bench.host_cpus = lambda nt=nt: nt
    run(f"oracle {nt} threads")
    bench.host_cpus = _orig
torch.set_num_threads(32)
bench.host_cpus = lambda: 64
run("torch 32 + oracle 64")
torch.set_num_threads(8)
run("torch 8 + oracle 64")
