"""Reference result of ONE fused frame on the CPU: eager binding getters (torch autograd) chained into the C oracle.

TEST INFRASTRUCTURE (oracle).  This is the data flow the reference runs for a `--bind_to_mesh` frame,
    update_mesh_properties  (scene/flame_gaussian_model.py:137-147)
    get_xyz / get_scaling / get_rotation / get_opacity / get_features  (scene/gaussian_model.py:113-160)
    GaussianRasterizer(...)  (gaussian_renderer/__init__.py:86-94; here oracle/splat_oracle.c)
    loss.backward()          (autograd through all of the above)
restated on CPU tensors, so that the fused CUDA op can be compared output by output: image, radii, the gradients of
the six raw parameter arrays, of `viewspace_points` and of the mesh vertices.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / parity_check legs may import this module.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from . import binding as ob
from . import rasterizer as orc

RAW = ("_xyz", "_rotation", "_scaling", "_opacity", "_features_dc", "_features_rest")


def fused_frame(params: Dict[str, torch.Tensor], verts: torch.Tensor, faces: torch.Tensor, cam, W: int, H: int,
                bg: torch.Tensor, sh_degree: int, dL_dimage: Optional[torch.Tensor] = None,
                scale_modifier: float = 1.0, threads: Optional[int] = None):
    """params: CPU float32 raw parameters + int `binding`; verts (V,3) posed mesh; cam: object with
    world_view_transform / full_proj_transform / camera_center / tanfovx / tanfovy.
    Returns dict(image (3,H,W) f32, radii (P,) i32, N, n_contrib_max, tile_list_max, final_T, and -- when
    `dL_dimage` is given -- grads {raw name: array}, 'means2D' (P,3), 'verts' (V,3))."""
    from collections import namedtuple

    if threads:
        torch.set_num_threads(threads)
    need = dL_dimage is not None
    leaves = {k: params[k].detach().float().clone().requires_grad_(need) for k in RAW}
    v = verts.detach().float().clone().requires_grad_(need)
    b = params["binding"].long()
    fr = ob.update_mesh_properties(v, faces)
    act = dict(
        means3D=ob.get_xyz(leaves["_xyz"], b, fr["face_center"], fr["face_orien_mat"], fr["face_scaling"]),
        scales=ob.get_scaling(leaves["_scaling"], b, fr["face_scaling"]),
        rotations=ob.get_rotation(leaves["_rotation"], b, fr["face_orien_quat"]),
        opacities=ob.get_opacity(leaves["_opacity"]),
        shs=ob.get_features(leaves["_features_dc"], leaves["_features_rest"]).contiguous())
    RS = namedtuple("RS", "image_height image_width tanfovx tanfovy bg scale_modifier viewmatrix projmatrix sh_degree "
                          "campos prefiltered debug")
    rs = RS(H, W, cam.tanfovx, cam.tanfovy, bg.float().cpu(), scale_modifier, cam.world_view_transform.cpu(),
            cam.full_proj_transform.cpu(), sh_degree, cam.camera_center.cpu(), False, False)
    Fn = orc.make_autograd_function()
    holder = {}

    class Keep(Fn):  # expose the oracle state (n_contrib, ranges) of this forward
        @staticmethod
        def forward(ctx, *a):
            out = Fn.forward(ctx, *a)
            holder["st"] = ctx.st
            return out

        @staticmethod
        def backward(ctx, *g):
            return Fn.backward(ctx, *g)

    P = leaves["_xyz"].shape[0]
    m2 = torch.zeros(P, 3, requires_grad=need)
    img, radii = Keep.apply(act["means3D"], m2, act["shs"], None, act["opacities"], act["scales"], act["rotations"],
                            None, rs)
    st = holder["st"]
    lens = st.ranges[:, 1].astype(np.int64) - st.ranges[:, 0].astype(np.int64)
    out = dict(image=img.detach().numpy(), radii=radii.numpy(), N=int(st.N), final_T=st.final_T,
               n_contrib_max=int(st.n_contrib.max()) if st.n_contrib.size else 0,
               tile_list_max=int(lens.max()) if lens.size else 0)
    if need:
        (img * dL_dimage.float().cpu()).sum().backward()
        g = {k: leaves[k].grad.numpy() for k in RAW}
        g["means2D"] = m2.grad.numpy()
        g["verts"] = v.grad.numpy()
        out["grads"] = g
    return out
