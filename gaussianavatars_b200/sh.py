"""Torch evaluation of real spherical harmonics up to degree 3 (the `convert_SHs_python` route of render();
reference twin: utils/sh_utils.py:57-112).  Table-driven: each basis function is (constant, polynomial in x,y,z)."""
import torch

_C0 = 0.28209479177387814
_C1 = 0.4886025119029199
_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
       1.445305721320277, -0.5900435899266435)


def sh_basis(deg: int, dirs: torch.Tensor) -> torch.Tensor:
    """dirs (...,3) unit -> (..., (deg+1)^2)."""
    x, y, z = dirs[..., 0], dirs[..., 1], dirs[..., 2]
    cols = [torch.full_like(x, _C0)]
    if deg > 0:
        cols += [-_C1 * y, _C1 * z, -_C1 * x]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        cols += [_C2[0] * xy, _C2[1] * yz, _C2[2] * (2.0 * zz - xx - yy), _C2[3] * xz, _C2[4] * (xx - yy)]
    if deg > 2:
        cols += [_C3[0] * y * (3 * xx - yy), _C3[1] * xy * z, _C3[2] * y * (4 * zz - xx - yy),
                 _C3[3] * z * (2 * zz - 3 * xx - 3 * yy), _C3[4] * x * (4 * zz - xx - yy), _C3[5] * z * (xx - yy),
                 _C3[6] * x * (xx - 3 * yy)]
    if deg > 3:
        raise NotImplementedError("SH degree > 3 is not used by GaussianAvatars (arguments/__init__.py:49)")
    return torch.stack(cols, dim=-1)


def eval_sh(deg: int, sh: torch.Tensor, dirs: torch.Tensor) -> torch.Tensor:
    """sh (..., C, K) with K >= (deg+1)^2, dirs (..., 3) -> (..., C)."""
    B = sh_basis(deg, dirs)
    return (sh[..., : B.shape[-1]] * B[..., None, :]).sum(dim=-1)
