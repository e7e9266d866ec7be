"""Drop-in for gaussian_renderer.render() (reference: gaussian_renderer/__init__.py:19-101).

`render(viewpoint_camera, pc, pipe, bg_color, scaling_modifier=1.0, override_color=None)` returns the same dict
{"render", "viewspace_points", "visibility_filter", "radii"}.  Two routes:

  * reference route  -- pc's getters (get_xyz / get_scaling / ...) feed `GaussianRasterizer` exactly as the
    reference does (also taken when pipe.compute_cov3D_python or pipe.convert_SHs_python is set);
  * fused route      -- when `pc` exposes the raw parameters (`_xyz, _rotation, _scaling, _opacity, _features_dc,
    _features_rest`) the binding of scene/gaussian_model.py:113-160 runs inside the preprocess kernel
    (`rasterize_bound`): no getter, no torch.cat, no (P,3,3) temporaries, one flat gradient buffer.
Camera matrices that live on the host are uploaded once and cached on the camera object (the reference re-uploads
three tensors per call, gaussian_renderer/__init__.py:44-47).
"""
from __future__ import annotations

import math

import torch

from .rasterizer import GaussianRasterizationSettings, GaussianRasterizer, hints_of, rasterize_bound, visible_of


def _camera_block(cam, device):
    cached = getattr(cam, "_gab200_dev", None)
    if cached is not None and cached[0] == device and cached[1] is cam.world_view_transform:
        return cached[2]
    blk = (cam.world_view_transform.to(device=device, dtype=torch.float32).contiguous(),
           cam.full_proj_transform.to(device=device, dtype=torch.float32).contiguous(),
           cam.camera_center.to(device=device, dtype=torch.float32).contiguous())
    try:
        cam._gab200_dev = (device, cam.world_view_transform, blk)
    except Exception:  # read-only camera objects: just do not cache
        pass
    return blk


def _settings(cam, pc, pipe, bg_color, scaling_modifier, device):
    view, proj, center = _camera_block(cam, device)
    return GaussianRasterizationSettings(
        image_height=int(cam.image_height), image_width=int(cam.image_width),
        tanfovx=math.tan(cam.FoVx * 0.5), tanfovy=math.tan(cam.FoVy * 0.5), bg=bg_color,
        scale_modifier=scaling_modifier, viewmatrix=view, projmatrix=proj, sh_degree=pc.active_sh_degree,
        campos=center, prefiltered=False, debug=bool(getattr(pipe, "debug", False)))


def _has_raw(pc):
    return all(hasattr(pc, n) for n in ("_xyz", "_rotation", "_scaling", "_opacity", "_features_dc", "_features_rest"))


def render_bound(viewpoint_camera, pc, pipe, bg_color, scaling_modifier=1.0, override_color=None):
    """Fused route (see module docstring)."""
    device = pc._xyz.device
    rs = _settings(viewpoint_camera, pc, pipe, bg_color, scaling_modifier, device)
    P = pc._xyz.shape[0]
    screenspace_points = torch.zeros((P, 3), dtype=pc._xyz.dtype, device=device, requires_grad=True)
    binding = getattr(pc, "binding", None)
    fc = fR = fs = None
    if binding is not None:
        if getattr(pc, "face_center", None) is None:
            pc.select_mesh_by_timestep(0)  # as the reference getters do (scene/gaussian_model.py:119-120)
        fc, fR, fs = pc.face_center, pc.face_orien_mat, pc.face_scaling
    rendered_image, radii = rasterize_bound(rs, pc._xyz, pc._rotation, pc._scaling, pc._opacity, pc._features_dc,
                                            pc._features_rest, binding, fc, fR, fs, means2D=screenspace_points,
                                            colors_precomp=override_color, grad_sink=pc)
    return {"render": rendered_image, "viewspace_points": screenspace_points, "visibility_filter": _visible(radii),
            "radii": radii}


def _visible(radii):
    """`radii > 0` (gaussian_renderer/__init__.py:100): the forward wrote it next to the radii (one byte per splat)."""
    v = visible_of(radii)
    return v if v is not None else radii > 0


def render(viewpoint_camera, pc, pipe, bg_color, scaling_modifier=1.0, override_color=None, fused=None):
    python_paths = bool(getattr(pipe, "compute_cov3D_python", False)) or bool(getattr(pipe, "convert_SHs_python", False))
    if fused is None:
        fused = _has_raw(pc) and not python_paths
    if fused:
        return render_bound(viewpoint_camera, pc, pipe, bg_color, scaling_modifier, override_color)

    # ---- reference route (data flow of gaussian_renderer/__init__.py:27-101) ----
    xyz = pc.get_xyz
    device = xyz.device
    screenspace_points = torch.zeros_like(xyz, dtype=xyz.dtype, requires_grad=True, device=device) + 0
    try:
        screenspace_points.retain_grad()
    except Exception:
        pass
    rs = _settings(viewpoint_camera, pc, pipe, bg_color, scaling_modifier, device)
    rasterizer = GaussianRasterizer(raster_settings=rs, hints=hints_of(pc))
    means3D, means2D, opacity = xyz, screenspace_points, pc.get_opacity
    scales = rotations = cov3D_precomp = None
    if getattr(pipe, "compute_cov3D_python", False):
        cov3D_precomp = pc.get_covariance(scaling_modifier)
    else:
        scales, rotations = pc.get_scaling, pc.get_rotation
    shs = colors_precomp = None
    if override_color is None:
        if getattr(pipe, "convert_SHs_python", False):
            from .sh import eval_sh
            shs_view = pc.get_features.transpose(1, 2).view(-1, 3, (pc.max_sh_degree + 1) ** 2)
            dir_pp = xyz - rs.campos.repeat(pc.get_features.shape[0], 1)
            dir_pp_normalized = dir_pp / dir_pp.norm(dim=1, keepdim=True)
            sh2rgb = eval_sh(pc.active_sh_degree, shs_view, dir_pp_normalized)
            colors_precomp = torch.clamp_min(sh2rgb + 0.5, 0.0)
        else:
            shs = pc.get_features
    else:
        colors_precomp = override_color
    rendered_image, radii = rasterizer(means3D=means3D, means2D=means2D, shs=shs, colors_precomp=colors_precomp,
                                       opacities=opacity, scales=scales, rotations=rotations,
                                       cov3D_precomp=cov3D_precomp)
    return {"render": rendered_image, "viewspace_points": screenspace_points, "visibility_filter": _visible(radii),
            "radii": radii}
