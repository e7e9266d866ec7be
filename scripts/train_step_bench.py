"""BASELINE.json config 3: bound avatar, 150k splats, 16 cameras, one --bind_to_mesh training step per camera:
face frame -> fused forward -> L1 (uint8 gt) -> backward (incl. dL/dverts) -> fused Adam on the six splat tensors.
Prints one JSON line per resolution."""
import json, os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
from gaussianavatars_b200 import synthetic as syn, l1_loss_u8
from gaussianavatars_b200.model import MeshBoundGaussians
from gaussianavatars_b200.renderer import render

dev = torch.device("cuda:0")
class Pipe: debug=False; compute_cov3D_python=False; convert_SHs_python=False
P = int(os.environ.get("P", 150000))
verts, faces = syn.head_mesh()
params = syn.avatar_splats(P, n_faces=faces.shape[0], seed=0, sh_degree=3)
bg = torch.ones(3, device=dev)
for (W, H) in ((550, 802), (1920, 1080)):
    pc = MeshBoundGaussians(params, 3, verts, faces, pose_fn=syn.pose_mesh, device=dev, requires_grad=True)
    opt = torch.optim.Adam(pc.parameters(), lr=1e-4, eps=1e-15, fused=True)
    cams = [syn.orbit_camera(W, H, azimuth_deg=-60 + 120 * (i + .5) / 16).to(dev) for i in range(16)]
    posed = [syn.pose_mesh(pc.verts_rest, i).contiguous().requires_grad_(True) for i in range(16)]
    gts = [torch.randint(0, 256, (3, H, W), dtype=torch.uint8, device=dev) for _ in range(2)]
    def step(i):
        opt.zero_grad(set_to_none=True)
        pc.update_mesh_properties(posed[i % 16])
        out = render(cams[i % 16], pc, Pipe, bg)
        loss = l1_loss_u8(out["render"], gts[i % 2])
        loss.backward()          # reaches the vertices through the face-frame kernel's backward
        opt.step()
    for i in range(8): step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 64
    e0.record()
    for i in range(K): step(i)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print(json.dumps({"config": "BASELINE configs[2]: 150k bound splats, 16 cameras, training step", "splats": P, "W": W, "H": H,
                      "ms_per_training_step": round(ms, 4), "steps_per_s": round(1e3 / ms, 1)}), flush=True)
