"""A few fused fwd+bwd frames of the headline workload (for ncu variant sweeps)."""
import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
from gaussianavatars_b200 import synthetic as syn, rasterizer as R
from gaussianavatars_b200.model import MeshBoundGaussians
from gaussianavatars_b200.renderer import render
dev = torch.device("cuda:0")
W, H, P = 1920, 1080, 100000
verts, faces = syn.head_mesh()
params = syn.avatar_splats(P, n_faces=faces.shape[0], seed=0, sh_degree=3)
cam = syn.orbit_camera(W, H).to(dev)
class Pipe: debug=False; compute_cov3D_python=False; convert_SHs_python=False
bg = torch.ones(3, device=dev)
gout = torch.randn(3, H, W, device=dev)
pc = MeshBoundGaussians(params, 3, verts, faces, device=dev, requires_grad=True)
pc.select_mesh_by_timestep(0)
for _ in range(int(os.environ.get("FRAMES", 3))):
    for p_ in pc.parameters(): p_.grad = None
    out = render(cam, pc, Pipe, bg)
    out["render"].backward(gout)
torch.cuda.synchronize()
