"""Dense, differentiable float64 model of the rasterizer (TEST INFRASTRUCTURE).

An independent anchor for the oracle's hand-restated backward (SURVEY.md 8c "independent correctness
anchors"): every pixel looks at every splat, order = stable sort by depth, tile membership enters
only as a boolean mask.  Gradients come from torch autograd, so they share no code with
oracle/splat_oracle.c's stage 4/5.  The module's deliberate non-analytic behaviours (Appendix B.4/B.5)
are modelled explicitly:
  - min(0.99, .) passes the gradient straight through;
  - the 1.3*tanfov guard band zeroes d/dt.x (d/dt.y) when clamped and ignores the clamp's t.z dependence;
  - quaternions are used unnormalised, no normalisation Jacobian;
  - dL/dscale is w.r.t. s = mod*scale (no extra `mod` factor) -> model scales as (mod*scale).detach()-shifted;
  - SH colours clamp at 0 with zero gradient where clamped.
Only suitable for small P and images (memory O(H*W*P)).
"""
from __future__ import annotations

import math

import torch

SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
SH_C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
         1.445305721320277, -0.5900435899266435]


def sh_basis(deg, d):
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    B = [torch.full_like(x, SH_C0)]
    if deg > 0:
        B += [-SH_C1 * y, SH_C1 * z, -SH_C1 * x]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        B += [SH_C2[0] * xy, SH_C2[1] * yz, SH_C2[2] * (2 * zz - xx - yy), SH_C2[3] * xz, SH_C2[4] * (xx - yy)]
    if deg > 2:
        B += [SH_C3[0] * y * (3 * xx - yy), SH_C3[1] * xy * z, SH_C3[2] * y * (4 * zz - xx - yy),
              SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy), SH_C3[4] * x * (4 * zz - xx - yy), SH_C3[5] * z * (xx - yy),
              SH_C3[6] * x * (xx - 3 * yy)]
    return torch.stack(B, dim=1)  # (P, nb)


def quat_to_R(q):
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    return torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=1).reshape(-1, 3, 3)


def render(means3D, means2D, opacities, viewmatrix, projmatrix, campos, W, H, tanfovx, tanfovy, bg, *, shs=None,
           sh_degree=0, colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None, scale_modifier=1.0,
           radii=None):
    """All tensor arguments float64.  `radii` (int, from the fp32 oracle) fixes the discrete tile rectangles so the
    comparison is not at the mercy of ceil() knife edges.  Returns (image (3,H,W), aux dict)."""
    dt = torch.float64
    P = means3D.shape[0]
    V = viewmatrix.reshape(16).to(dt)
    Pm = projmatrix.reshape(16).to(dt)
    fx, fy = W / (2.0 * tanfovx), H / (2.0 * tanfovy)
    x, y, z = means3D[:, 0], means3D[:, 1], means3D[:, 2]
    tx = V[0] * x + V[4] * y + V[8] * z + V[12]
    ty = V[1] * x + V[5] * y + V[9] * z + V[13]
    tz = V[2] * x + V[6] * y + V[10] * z + V[14]
    hx = Pm[0] * x + Pm[4] * y + Pm[8] * z + Pm[12]
    hy = Pm[1] * x + Pm[5] * y + Pm[9] * z + Pm[13]
    hw = Pm[3] * x + Pm[7] * y + Pm[11] * z + Pm[15]
    p_w = 1.0 / (hw + 0.0000001)
    ndc_x, ndc_y = hx * p_w, hy * p_w
    in_front = tz > 0.2

    if cov3D_precomp is not None:
        c3 = cov3D_precomp
        Sigma = torch.stack([c3[:, 0], c3[:, 1], c3[:, 2], c3[:, 1], c3[:, 3], c3[:, 4], c3[:, 2], c3[:, 4], c3[:, 5]],
                            dim=1).reshape(P, 3, 3)
    else:
        R = quat_to_R(rotations)
        # gradient w.r.t. s=mod*scale reported as dL/dscale: value mod*scale, derivative 1
        s = scales + ((scale_modifier - 1.0) * scales).detach()
        Sigma = R @ torch.diag_embed(s * s) @ R.transpose(1, 2)

    limx, limy = 1.3 * tanfovx, 1.3 * tanfovy
    txtz, tytz = tx / tz, ty / tz
    cx = (txtz < -limx) | (txtz > limx)
    cy = (tytz < -limy) | (tytz > limy)
    txc = torch.where(cx, (txtz.clamp(-limx, limx) * tz).detach(), tx)
    tyc = torch.where(cy, (tytz.clamp(-limy, limy) * tz).detach(), ty)
    zero = torch.zeros_like(tz)
    J = torch.stack([fx / tz, zero, -(fx * txc) / (tz * tz), zero, fy / tz, -(fy * tyc) / (tz * tz)], dim=1).reshape(P, 2, 3)
    Wv = torch.stack([V[0], V[4], V[8], V[1], V[5], V[9], V[2], V[6], V[10]]).reshape(3, 3)
    T = J @ Wv
    cov = T @ Sigma @ T.transpose(1, 2)
    a, b, c = cov[:, 0, 0] + 0.3, cov[:, 0, 1], cov[:, 1, 1] + 0.3
    det = a * c - b * b
    # conic with the 1/(det^2+1e-7)-style guard irrelevant at fp64 tolerance
    conA, conB, conC = c / det, -b / det, a / det

    px = ((ndc_x + 1.0) * W - 1.0) * 0.5 + means2D[:, 0] * (0.5 * W)
    py = ((ndc_y + 1.0) * H - 1.0) * 0.5 + means2D[:, 1] * (0.5 * H)

    if colors_precomp is not None:
        rgb = colors_precomp
    else:
        d = means3D - campos.reshape(1, 3)
        d = d / d.norm(dim=1, keepdim=True)
        B = sh_basis(sh_degree, d)
        nb = B.shape[1]
        rgb = (B[:, :, None] * shs[:, :nb, :]).sum(dim=1) + 0.5
        rgb = torch.clamp_min(rgb, 0.0)

    # discrete tile rectangles
    gx, gy = (W + 15) // 16, (H + 15) // 16
    if radii is None:
        mid = 0.5 * (a + c)
        lam = mid + torch.sqrt(torch.clamp_min(mid * mid - det, 0.1))
        radii = torch.ceil(3.0 * torch.sqrt(lam)).to(torch.int64)
    rad = radii.to(dt)
    pxd, pyd = px.detach(), py.detach()
    x0 = torch.clamp(((pxd - rad) / 16).trunc(), 0, gx)
    y0 = torch.clamp(((pyd - rad) / 16).trunc(), 0, gy)
    x1 = torch.clamp(((pxd + rad + 15) / 16).trunc(), 0, gx)
    y1 = torch.clamp(((pyd + rad + 15) / 16).trunc(), 0, gy)
    visible = in_front & (radii > 0) & ((x1 - x0) * (y1 - y0) > 0)

    order = torch.argsort(tz.detach().to(torch.float32), stable=True)  # fp32 depth bits decide, ties by id
    ys, xs = torch.meshgrid(torch.arange(H, dtype=dt), torch.arange(W, dtype=dt), indexing="ij")
    pixx, pixy = xs.reshape(-1, 1), ys.reshape(-1, 1)  # (HW,1)
    tilex, tiley = (pixx / 16).floor(), (pixy / 16).floor()

    def g(v):
        return v[order][None, :]

    member = g(visible) & (tilex >= g(x0)) & (tilex < g(x1)) & (tiley >= g(y0)) & (tiley < g(y1))
    dx, dy = g(px) - pixx, g(py) - pixy
    power = -0.5 * (g(conA) * dx * dx + g(conC) * dy * dy) - g(conB) * dx * dy
    Gs = torch.exp(torch.clamp_max(power, 0.0))
    og = g(opacities.reshape(-1)) * Gs
    alpha = og + (torch.clamp_max(og, 0.99) - og).detach()
    valid = member & (power <= 0) & (alpha.detach() >= 1.0 / 255.0)
    aeff = torch.where(valid, alpha, torch.zeros_like(alpha))
    one_minus = 1.0 - aeff
    T_incl = torch.cumprod(one_minus, dim=1)
    T_before = torch.cat((torch.ones_like(T_incl[:, :1]), T_incl[:, :-1]), dim=1)
    stop = valid & (T_incl.detach() < 0.0001)
    stopped = torch.cumsum(stop.to(torch.int64), dim=1) > 0  # inclusive: the stopping instance itself is dropped
    keep = valid & ~stopped
    w = torch.where(keep, aeff * T_before, torch.zeros_like(aeff))
    C = w @ rgb[order]  # (HW,3)
    T_final = torch.where(keep, one_minus, torch.ones_like(one_minus)).prod(dim=1)
    out = C + T_final[:, None] * bg.reshape(1, 3)
    img = out.t().reshape(3, H, W)
    return img, dict(radii=radii, visible=visible, T_final=T_final.reshape(H, W), n_keep=keep.sum(dim=1).reshape(H, W))
