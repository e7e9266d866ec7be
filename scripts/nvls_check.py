"""N-rank check (torchrun): gradients reduced by the fused NVLS path == NCCL all-reduce of per-rank gradients."""
import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch, torch.distributed as dist
from gaussianavatars_b200 import synthetic as syn, dist as gdist
from gaussianavatars_b200.model import MeshBoundGaussians
from gaussianavatars_b200.renderer import render
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
class Pipe: debug=False; compute_cov3D_python=False; convert_SHs_python=False
verts, faces = syn.head_mesh(n_lat=20, n_lon=36)
params = syn.avatar_splats(20000, n_faces=faces.shape[0], seed=0, sh_degree=3, scale_gain=2.0)
pc = MeshBoundGaussians(params, 3, verts, faces, device=dev, requires_grad=True)
pc.select_mesh_by_timestep(0)
cam = syn.orbit_camera(640, 480, azimuth_deg=-30 + 60 * rank / max(world - 1, 1)).to(dev)
bg = torch.ones(3, device=dev)
gout = torch.randn(3, 480, 640, generator=torch.Generator().manual_seed(1)).to(dev)
# reference: local grads + NCCL
render(cam, pc, Pipe, bg)["render"].backward(gout)
gdist.allreduce_splat_grads(pc)
ref = [p.grad.clone() for p in pc.parameters()]
for p in pc.parameters(): p.grad = None
symm = gdist.SymmetricGradBuffer(pc)
print(f"rank {rank}: nvls enabled={symm.enabled} {getattr(symm, 'error', '')}", flush=True)
if symm.enabled:
    pc.symm_grad = symm
    for it in range(3):
        for p in pc.parameters(): p.grad = None
        symm.begin()
        render(cam, pc, Pipe, bg)["render"].backward(gout)
        symm.end()
    torch.cuda.synchronize()
    worst = 0.0
    for p, r in zip(pc.parameters(), ref):
        err = float((p.grad - r).abs().max() / (r.abs().max() + 1e-30)); worst = max(worst, err)
    print(f"rank {rank}: max rel diff NVLS vs NCCL = {worst:.3e}", flush=True)
    assert worst < 1e-4
dist.barrier(); dist.destroy_process_group()
