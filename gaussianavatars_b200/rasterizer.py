"""Host-side mirror of the reference operator surface, backed by the sm_100a C-ABI library.

Drop-in names (the only two GaussianAvatars imports, gaussian_renderer/__init__.py:15):
    GaussianRasterizationSettings, GaussianRasterizer
with the semantics of diff_gaussian_rasterization/__init__.py of the pinned submodule:
same argument names, same "exactly one of" exceptions, same outputs (color (3,H,W), radii (P,) int32), same
gradient tuple (means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, None).

New, opt-in fused surface (SURVEY.md 8b "Fused surface"): `rasterize_bound(...)` takes the RAW GaussianModel
parameters plus the per-face frame and runs scene/gaussian_model.py:113-160 inside the preprocess kernel; its
backward returns gradients for the raw parameters and for face_center / face_orien_mat / face_scaling.
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import NamedTuple, Optional

import torch
import torch.nn as nn

from . import _native as N


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


# Binning policy of the library (include/gab200_rasterizer.h `exact_binning`).  False (default) drops (splat, tile)
# pairs that provably contribute nothing; image, radii and gradients are unchanged.  True reproduces the
# reference's full 3-sigma bounding-square instance list (used by the key/sort parity tests).
_EXACT_BINNING = False


def set_exact_binning(flag: bool):
    global _EXACT_BINNING
    _EXACT_BINNING = bool(flag)


def _f32c(t: Optional[torch.Tensor], name: str, device):
    if t is None or t.numel() == 0:
        return None
    if t.device != device:
        raise ValueError(f"{name} must live on {device}, got {t.device}")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _cam(t: torch.Tensor, name: str, device):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a tensor")
    if t.device != device or t.dtype != torch.float32:
        t = t.to(device=device, dtype=torch.float32)
    return t if t.is_contiguous() else t.contiguous()


class FrameHints:
    """What one caller (one splat model) has learnt about its frames: per (device, W, H, P), the instance count of the
    last frame (the capacity the next one runs with: gab200_forward_args.binning_hint) and the depth-key range
    (depth_hint_*).  Hints never change a result, only how much of the forward is enqueued before the host learns N.
    `render()` keeps one on the model object (`pc._gab200_hints`); callers of the bare operator surface that cannot
    pass one (the reference's own render() builds a new GaussianRasterizer per frame) share `_default_hints`."""

    MAX_SHAPES = 64  # P changes at every densification: do not let the per-shape hints pile up

    def __init__(self):
        self.shapes = {}
        self.seq = 0
        self.last = None  # info dict of the last forward that used these hints

    def get(self, key):
        return self.shapes.get(key, (0, (0, 0)))

    def set_depth(self, key, depth_range):
        self.shapes[key] = (self.get(key)[0], tuple(depth_range))

    def set_capacity(self, key, capacity: int):
        self.shapes[key] = (int(capacity), self.get(key)[1])

    def put(self, key, n: int, depth_range):
        if len(self.shapes) >= self.MAX_SHAPES and key not in self.shapes:
            self.shapes.clear()
        self.shapes[key] = (min(int(n * 1.25) + 4096, 2**31 - 1), depth_range)


_default_hints = FrameHints()
_SYNC_POLICY = "late"   # "late": sync-free enqueue + end-of-call check whenever a capacity hint exists; "exact": always mid-frame


def set_sync_policy(policy: str):
    """"late" (default) or "exact" -- see include/gab200_rasterizer.h gab200_sync_mode.  Results are identical."""
    global _SYNC_POLICY
    if policy not in ("late", "exact"):
        raise ValueError("policy must be 'late' or 'exact'")
    _SYNC_POLICY = policy


def hints_of(obj) -> "FrameHints":
    """The FrameHints attached to a model object (created on first use); `_default_hints` if it cannot carry one."""
    h = getattr(obj, "_gab200_hints", None)
    if h is None:
        h = FrameHints()
        try:
            obj._gab200_hints = h
        except Exception:
            return _default_hints
    return h


def _fill_common(a: N.ForwardArgs, rs: GaussianRasterizationSettings, device, P: int, need_backward: bool):
    a.abi_version = N.ABI_VERSION
    a.P = P
    a.sh_degree = int(rs.sh_degree)
    a.image_width = int(rs.image_width)
    a.image_height = int(rs.image_height)
    a.tanfovx = float(rs.tanfovx)
    a.tanfovy = float(rs.tanfovy)
    a.scale_modifier = float(rs.scale_modifier)
    a.prefiltered = int(bool(rs.prefiltered))
    a.debug = int(bool(rs.debug))
    a.need_backward = int(need_backward)
    a.exact_binning = int(_EXACT_BINNING)
    cams = (_cam(rs.bg, "bg", device), _cam(rs.viewmatrix, "viewmatrix", device),
            _cam(rs.projmatrix, "projmatrix", device), _cam(rs.campos, "campos", device))
    a.bg, a.viewmatrix, a.projmatrix, a.campos = (t.data_ptr() for t in cams)
    return cams


_KEEP_LAST = False
_last = None
_last_info = {}


def _widen_depth_range(kmin: int, kmax: int):
    """The last frame's visible depth-key range plus 1/8 of its width either side (keys are fp32 bit patterns of
    positive depths: monotonic, so a range in key space is a range in depth).  A frame that falls outside is still
    sorted correctly -- outliers share the two end buckets -- so the margin only has to keep that rare."""
    pad = max((kmax - kmin) // 8, 1 << 12)
    return max(kmin - pad, 1), min(kmax + pad, 0xFFFFFFFE)


def keep_last_state(flag: bool):
    """Parity/debug hook: retain the last forward's scratch so `export_last_binning()` can read the sorted stream."""
    global _KEEP_LAST, _last
    _KEEP_LAST = bool(flag)
    if not flag:
        _last = None


def export_last_binning():
    """(keys u64 as int64 tensor, values int32 tensor, ranges (tiles,2) int32 tensor, num_rendered) of the last forward."""
    if _last is None:
        raise RuntimeError("keep_last_state(True) was not set before the forward")
    a, st, holder, device = _last
    n = int(st.num_rendered)
    gx, gy = (a.image_width + 15) // 16, (a.image_height + 15) // 16
    keys = torch.empty((max(n, 1),), dtype=torch.int64, device=device)
    vals = torch.empty((max(n, 1),), dtype=torch.int32, device=device)
    ranges = torch.empty((gx * gy, 2), dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        stream = torch.cuda.current_stream(device).cuda_stream
        N.check(N.lib().gab200_export_binning(C.byref(a), C.byref(st), keys.data_ptr(), vals.data_ptr(),
                                              ranges.data_ptr(), C.c_void_p(stream)), "gab200_export_binning")
    return keys[:n], vals[:n], ranges, n


def last_frame_info() -> dict:
    """How the last forward on this process went: instances, capacity, depth-sort path, sync mode, attempts."""
    return dict(_last_info)


class CaptureSlot:
    """What a forward needs while its stream is being captured into a CUDA graph (GAB200_SYNC_NONE): a fixed
    capacity, a pinned host copy of the frame counters, and a sticky device flag the library raises when a replay
    overflows the capacity (graph.py owns one per captured step)."""

    def __init__(self, device, capacity: int, depth_range=(0, 0)):
        self.capacity = int(capacity)
        self.depth_range = depth_range
        self.counters = torch.zeros(N.NUM_COUNTERS, dtype=torch.int32).pin_memory()
        self.flag = torch.zeros(1, dtype=torch.int32, device=device)   # sticky: set by the library on overflow
        self.flag_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.seq = 1
        self.info = {}


_capture_slot = None
_tls = threading.local()


def visible_of(radii: torch.Tensor):
    """The `radii > 0` mask the forward that produced `radii` wrote beside it (None if that was not the last forward
    of this thread)."""
    v = getattr(_tls, "visible", None)
    return v[1] if v is not None and v[0] == radii.data_ptr() else None


def _run_forward(a: N.ForwardArgs, device, need_backward: bool, hints: Optional[FrameHints] = None):
    global _last, _last_info
    H, W, P = a.image_height, a.image_width, a.P
    color = torch.empty((3, H, W), dtype=torch.float32, device=device)
    radii = torch.empty((P,), dtype=torch.int32, device=device)
    visible = torch.empty((P,), dtype=torch.bool, device=device)   # radii > 0, written by the preprocess kernel
    a.out_color, a.radii, a.visibility = color.data_ptr(), radii.data_ptr(), visible.data_ptr()
    _tls.visible = (radii.data_ptr(), visible)   # renderer.py hands it out as `visibility_filter`
    cb, holder = N.begin_forward(device, need_backward)
    a.alloc_geom = a.alloc_binning = a.alloc_image = cb
    st = N.FrameState()
    key = (device, W, H, P)
    slot = _capture_slot
    if slot is not None:      # graph capture: fixed capacity, no host wait; graph.py reads slot.counters after replays
        a.sync_mode = N.SYNC_NONE
        a.binning_hint = slot.capacity
        a.depth_hint_lo, a.depth_hint_hi = slot.depth_range
        a.frame_seq = slot.seq
        a.counters_host = slot.counters.data_ptr()
        a.overflow_flag = slot.flag.data_ptr()
    else:
        hints = hints if hints is not None else _default_hints
        a.binning_hint, (a.depth_hint_lo, a.depth_hint_hi) = hints.get(key)
        a.sync_mode = N.SYNC_LATE if (_SYNC_POLICY == "late" and a.binning_hint > 0) else N.SYNC_EXACT
        hints.seq = (hints.seq + 1) & 0x7FFFFFFF
        a.frame_seq = hints.seq
    with torch.cuda.device(device):
        stream = torch.cuda.current_stream(device).cuda_stream
        n = N.lib().gab200_forward(C.byref(a), C.byref(st), C.c_void_p(stream))
    N.check(n, "gab200_forward")
    info = dict(num_rendered=int(st.num_rendered), capacity=int(st.binning_capacity), sync_mode=int(a.sync_mode),
                depth_sort_path=int(st.depth_sort_path), attempts=int(st.attempts))
    if slot is not None:
        slot.info = info
    else:
        hints.put(key, n, _widen_depth_range(st.depth_key_min, st.depth_key_max)
                  if st.depth_key_min <= st.depth_key_max else (0, 0))
        hints.last = info
    _last_info = info
    if _KEEP_LAST:
        # pooled (inference) scratch stays valid until the next no_grad forward on this device
        _last = (a, st, holder, device)
    return color, radii, st, holder


class _RasterizeGaussians(torch.autograd.Function):
    """Reference surface (ACTIVATED inputs)."""

    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                raster_settings):
        rs = raster_settings
        device = means3D.device
        if device.type != "cuda":
            raise RuntimeError("gaussianavatars_b200 has no CPU path: tensors must be CUDA tensors")
        if means3D.ndim != 2 or means3D.shape[1] != 3:
            raise RuntimeError("means3D must have dimensions (num_points, 3)")
        P = means3D.shape[0]
        need_bw = any(ctx.needs_input_grad)
        a = N.ForwardArgs()
        cams = _fill_common(a, rs, device, P, need_bw)
        a.input_mode = N.INPUT_ACTIVATED
        means3D = _f32c(means3D, "means3D", device)
        sh = _f32c(sh, "shs", device)
        colors_precomp = _f32c(colors_precomp, "colors_precomp", device)
        opacities = _f32c(opacities, "opacities", device)
        scales = _f32c(scales, "scales", device)
        rotations = _f32c(rotations, "rotations", device)
        cov3Ds_precomp = _f32c(cov3Ds_precomp, "cov3D_precomp", device)
        a.sh_coeffs = 0 if sh is None else sh.shape[1]
        a.means3D, a.opacities = N.ptr(means3D), N.ptr(opacities)
        a.scales, a.rotations, a.cov3D_precomp = N.ptr(scales), N.ptr(rotations), N.ptr(cov3Ds_precomp)
        a.shs, a.colors_precomp = N.ptr(sh), N.ptr(colors_precomp)
        color, radii, st, holder = _run_forward(a, device, need_bw, _hints_for_next_call())
        if need_bw:
            ctx.args, ctx.state, ctx.holder = a, st, holder
            ctx.keep = (cams, means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, radii)
            ctx.M = a.sh_coeffs
        ctx.mark_non_differentiable(radii)
        return color, radii

    @staticmethod
    def backward(ctx, grad_out_color, _grad_radii):
        a, st = ctx.args, ctx.state
        cams, means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, radii = ctx.keep
        device = means3D.device
        P, M = a.P, ctx.M
        g = grad_out_color if grad_out_color.is_contiguous() else grad_out_color.contiguous()
        e = lambda *s: torch.empty(s, dtype=torch.float32, device=device)  # noqa: E731
        d_means3D, d_means2D, d_opac = e(P, 3), e(P, 3), e(P, 1)
        d_colors = e(P, 3)
        d_sh = e(P, M, 3) if sh is not None else None
        d_scales = e(P, 3) if scales is not None else None
        d_rots = e(P, 4) if rotations is not None else None
        d_cov = e(P, 6)
        b = N.BackwardArgs()
        b.abi_version = N.ABI_VERSION
        b.fwd, b.state = C.pointer(a), C.pointer(st)
        b.dL_dout_color = g.data_ptr()
        b.dL_dmeans3D, b.dL_dmeans2D, b.dL_dopacity = d_means3D.data_ptr(), d_means2D.data_ptr(), d_opac.data_ptr()
        b.dL_dcolors, b.dL_dshs = d_colors.data_ptr(), N.ptr(d_sh)
        b.dL_dscales, b.dL_drotations, b.dL_dcov3D = N.ptr(d_scales), N.ptr(d_rots), d_cov.data_ptr()
        with torch.cuda.device(device):
            stream = torch.cuda.current_stream(device).cuda_stream
            N.check(N.lib().gab200_backward(C.byref(b), C.c_void_p(stream)), "gab200_backward")
        ctx.holder = None
        return (d_means3D, d_means2D, d_sh, d_colors if colors_precomp is not None else None, d_opac, d_scales, d_rots,
                d_cov if cov3Ds_precomp is not None else None, None)


_next_hints = None


def _hints_for_next_call():
    global _next_hints
    h, _next_hints = _next_hints, None
    return h


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                        raster_settings, hints: Optional[FrameHints] = None):
    global _next_hints
    _next_hints = hints  # autograd.Function.apply takes tensors and plain values; the hints object rides beside it
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                     cov3Ds_precomp, raster_settings)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings: GaussianRasterizationSettings, hints: Optional[FrameHints] = None):
        """`hints` is an extension over the reference's constructor (optional; see FrameHints)."""
        super().__init__()
        self.raster_settings = raster_settings
        self.hints = hints

    def markVisible(self, positions):
        with torch.no_grad():
            rs = self.raster_settings
            device = positions.device
            pos = _f32c(positions, "positions", device)
            out = torch.empty((pos.shape[0],), dtype=torch.uint8, device=device)
            view, proj = _cam(rs.viewmatrix, "viewmatrix", device), _cam(rs.projmatrix, "projmatrix", device)
            with torch.cuda.device(device):
                stream = torch.cuda.current_stream(device).cuda_stream
                N.check(N.lib().gab200_mark_visible(pos.shape[0], pos.data_ptr(), view.data_ptr(), proj.data_ptr(),
                                                    out.data_ptr(), C.c_void_p(stream)), "gab200_mark_visible")
        return out.bool()

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations,
                                   cov3D_precomp, self.raster_settings, self.hints)


# ================================================================================================================
# Fused surface
# ================================================================================================================
_face_csr_cache = {}   # id(binding tensor the caller passed) -> (that tensor, its _version, F, int32 copy, CSR tuple)


def _face_csr(binding: torch.Tensor, num_faces: int, chunk: int = 16):
    """(int32 contiguous binding, face-sorted view of it for the backward's per-face reduction
    (gab200_backward_args.face_*)).  The binding only changes at densification
    (scene/gaussian_model.py:472-474,495-497), so this runs once per change.  Keyed on the caller's OWN tensor (a
    reference is kept, so its id cannot be recycled) and its in-place version: a temporary `.to(int32)` copy whose
    address the allocator reuses can never alias another model's entry."""
    key = id(binding)
    hit = _face_csr_cache.get(key)
    if hit is not None and hit[0] is binding and hit[1] == binding._version and hit[2] == num_faces:
        return hit[3], hit[4]
    b32 = binding if binding.dtype == torch.int32 and binding.is_contiguous() else binding.to(torch.int32).contiguous()
    b = b32.long()
    perm = torch.argsort(b, stable=True).to(torch.int32)
    counts = torch.bincount(b, minlength=num_faces)
    starts = torch.cumsum(counts, 0) - counts
    nchunks = (counts + chunk - 1) // chunk
    face = torch.repeat_interleave(torch.arange(num_faces, device=b.device), nchunks)
    first = torch.cumsum(nchunks, 0) - nchunks
    within = torch.arange(face.shape[0], device=b.device) - first[face]
    c_start = starts[face] + within * chunk
    c_end = torch.minimum(c_start + chunk, starts[face] + counts[face])
    out = (perm.contiguous(), face.to(torch.int32).contiguous(), c_start.to(torch.int32).contiguous(),
           c_end.to(torch.int32).contiguous())
    if len(_face_csr_cache) >= 8:
        _face_csr_cache.clear()
    _face_csr_cache[key] = (binding, binding._version, num_faces, b32, out)
    return b32, out


class _RasterizeBound(torch.autograd.Function):
    @staticmethod
    def forward(ctx, _xyz, means2D, _rotation, _scaling, _opacity, f_dc, f_rest, face_center, face_orien_mat,
                face_scaling, binding, colors_precomp, raster_settings, grad_sink=None):
        rs = raster_settings
        ctx.grad_sink = grad_sink
        device = _xyz.device
        if device.type != "cuda":
            raise RuntimeError("gaussianavatars_b200 has no CPU path: tensors must be CUDA tensors")
        P = _xyz.shape[0]
        need_bw = any(ctx.needs_input_grad)
        a = N.ForwardArgs()
        cams = _fill_common(a, rs, device, P, need_bw)
        a.input_mode = N.INPUT_BOUND_RAW
        _xyz, _rotation = _f32c(_xyz, "_xyz", device), _f32c(_rotation, "_rotation", device)
        _scaling, _opacity = _f32c(_scaling, "_scaling", device), _f32c(_opacity, "_opacity", device)
        f_dc, f_rest = _f32c(f_dc, "_features_dc", device), _f32c(f_rest, "_features_rest", device)
        colors_precomp = _f32c(colors_precomp, "colors_precomp", device)
        M = 1 + (0 if f_rest is None else f_rest.shape[1])
        a.sh_coeffs = M
        a.means3D, a.rotations, a.scales, a.opacities = _xyz.data_ptr(), _rotation.data_ptr(), _scaling.data_ptr(), \
            _opacity.data_ptr()
        a.sh_dc, a.sh_rest, a.colors_precomp = N.ptr(f_dc), N.ptr(f_rest), N.ptr(colors_precomp)
        F = 0
        binding_orig = binding
        if binding is not None:
            if binding.dtype != torch.int32 or not binding.is_contiguous():
                binding = _face_csr(binding_orig, face_center.shape[0])[0]  # converted once per binding, not per frame
            face_center = _f32c(face_center, "face_center", device)
            face_orien_mat = _f32c(face_orien_mat, "face_orien_mat", device)
            face_scaling = _f32c(face_scaling, "face_scaling", device)
            F = face_center.shape[0]
            a.binding, a.num_faces = binding.data_ptr(), F
            a.face_center, a.face_orien_mat, a.face_scaling = face_center.data_ptr(), face_orien_mat.data_ptr(), \
                face_scaling.data_ptr()
        color, radii, st, holder = _run_forward(a, device, need_bw,
                                                hints_of(grad_sink) if grad_sink is not None else None)
        if need_bw:
            ctx.args, ctx.state, ctx.holder = a, st, holder
            ctx.keep = (cams, _xyz, _rotation, _scaling, _opacity, f_dc, f_rest, face_center, face_orien_mat,
                        face_scaling, binding, colors_precomp, radii)
            ctx.dims = (P, M, F)
            ctx.face_shapes = None if binding is None else (face_center.shape, face_orien_mat.shape,
                                                            face_scaling.shape)
            # face-frame gradients are only produced when something upstream of the frame trains (FLAME parameters)
            ctx.want_face = binding is not None and any(ctx.needs_input_grad[7:10])
            ctx.csr = _face_csr(binding_orig, F)[1] if ctx.want_face else None
        ctx.mark_non_differentiable(radii)
        return color, radii

    @staticmethod
    def backward(ctx, grad_out_color, _grad_radii):
        a, st = ctx.args, ctx.state
        P, M, F = ctx.dims
        device = ctx.keep[1].device
        binding, colors_precomp = ctx.keep[10], ctx.keep[11]
        g = grad_out_color if grad_out_color.is_contiguous() else grad_out_color.contiguous()
        # one flat buffer for all per-splat parameter gradients (dist.py all-reduces it in ONE collective):
        # [_xyz 3 | _rotation 4 | _scaling 3 | _opacity 1 | f_dc 3 | f_rest 3(M-1)]  = 59 floats/splat at SH3
        widths = (3, 4, 3, 1, 3, 3 * (M - 1))
        # frame-sharded data parallel with NVLS: the gradients are reduced INTO the symmetric buffer by the kernel
        symm = getattr(ctx.grad_sink, "symm_grad", None) if ctx.grad_sink is not None else None
        use_symm = symm is not None and symm.enabled and symm.numel == P * sum(widths) and colors_precomp is None
        # "push": the kernel reduces into every replica with multimem.red; "two_shot": plain stores into the local
        # replica, reduced afterwards by the NVLS all-reduce kernel (dist.SymmetricGradBuffer.end)
        use_mc = use_symm and getattr(symm, "mode", "push") == "push"
        flat = symm.flat if use_symm else torch.empty((P * sum(widths),), dtype=torch.float32, device=device)
        views, off = [], 0
        for w in widths:
            views.append(flat[off:off + P * w])
            off += P * w
        d_xyz, d_rot, d_scale, d_opac = views[0].view(P, 3), views[1].view(P, 4), views[2].view(P, 3), views[3].view(P, 1)
        d_dc = views[4].view(P, 1, 3)
        d_rest = views[5].view(P, M - 1, 3) if M > 1 else None
        d_means2D = torch.empty((P, 3), dtype=torch.float32, device=device)
        d_colors = torch.empty((P, 3), dtype=torch.float32, device=device) if colors_precomp is not None else None
        d_fc = d_fR = d_fs = None
        if ctx.want_face:
            fshape = ctx.face_shapes
            d_fc = torch.empty(fshape[0], dtype=torch.float32, device=device)
            d_fR = torch.empty(fshape[1], dtype=torch.float32, device=device)
            d_fs = torch.empty(fshape[2], dtype=torch.float32, device=device)
        b = N.BackwardArgs()
        b.abi_version = N.ABI_VERSION
        b.fwd, b.state = C.pointer(a), C.pointer(st)
        b.dL_dout_color = g.data_ptr()
        # parameter-gradient destinations: local addresses, or the same offsets inside the NVLS multicast mapping
        base = (symm.mc_ptr - flat.data_ptr()) if use_mc else 0
        b.grads_are_multicast = int(use_mc)
        b.dL_dmeans3D, b.dL_dopacity = d_xyz.data_ptr() + base, d_opac.data_ptr() + base
        b.dL_dmeans2D = d_means2D.data_ptr()
        b.dL_dcolors = N.ptr(d_colors)
        b.dL_dsh_dc = d_dc.data_ptr() + base
        b.dL_dsh_rest = None if d_rest is None else d_rest.data_ptr() + base
        b.dL_dscales, b.dL_drotations = d_scale.data_ptr() + base, d_rot.data_ptr() + base
        b.dL_dface_center, b.dL_dface_orien_mat, b.dL_dface_scaling = N.ptr(d_fc), N.ptr(d_fR), N.ptr(d_fs)
        if ctx.csr is not None:
            perm, c_face, c_start, c_end = ctx.csr
            b.face_perm, b.face_chunk_face = perm.data_ptr(), c_face.data_ptr()
            b.face_chunk_start, b.face_chunk_end = c_start.data_ptr(), c_end.data_ptr()
            b.num_face_chunks = c_face.shape[0]
        with torch.cuda.device(device):
            stream = torch.cuda.current_stream(device).cuda_stream
            N.check(N.lib().gab200_backward(C.byref(b), C.c_void_p(stream)), "gab200_backward")
        ctx.holder = None
        if ctx.grad_sink is not None:  # dist.py: ONE all-reduce over this buffer instead of six
            ctx.grad_sink.flat_grad = flat
            ctx.grad_sink._gab200_mc_used = bool(use_symm)  # SymmetricGradBuffer.end() only trusts the replica if set
        return (d_xyz, d_means2D, d_rot, d_scale, d_opac, d_dc, d_rest, d_fc, d_fR, d_fs, None, d_colors, None, None)


def rasterize_bound(raster_settings: GaussianRasterizationSettings, _xyz, _rotation, _scaling, _opacity,
                    features_dc, features_rest, binding=None, face_center=None, face_orien_mat=None,
                    face_scaling=None, means2D=None, colors_precomp=None, grad_sink=None):
    """Fused binding + rasterization.  Returns (color (3,H,W), radii (P,) int32).

    binding=None is the identity frame (a plain GaussianModel, scene/gaussian_model.py:115-116,127-128,142-143).
    `means2D` is the usual (P,3) gradient holder (its .grad receives dL/dmean2D in NDC units)."""
    if means2D is None:
        means2D = torch.zeros((_xyz.shape[0], 3), dtype=torch.float32, device=_xyz.device)
    if _opacity.ndim == 1:
        _opacity = _opacity[:, None]
    return _RasterizeBound.apply(_xyz, means2D, _rotation, _scaling, _opacity, features_dc, features_rest,
                                 face_center, face_orien_mat, face_scaling, binding, colors_precomp, raster_settings,
                                 grad_sink)


def bind_activate(raster_settings_or_modifier, _xyz, _rotation, _scaling, _opacity, binding=None, face_center=None,
                  face_orien_mat=None, face_scaling=None):
    """Exports what the fused preprocess computes for the binding (no autograd): world means3D (P,3), opacities (P,1),
    scales (P,3), cov3D (P,6).  Same device code as the fused forward -> bit-identical values."""
    device = _xyz.device
    P = _xyz.shape[0]
    mod = raster_settings_or_modifier.scale_modifier if isinstance(raster_settings_or_modifier, tuple) \
        else float(raster_settings_or_modifier)
    a = N.ForwardArgs()
    a.abi_version, a.input_mode, a.P, a.scale_modifier = N.ABI_VERSION, N.INPUT_BOUND_RAW, P, mod
    keep = [_f32c(_xyz.detach(), "_xyz", device), _f32c(_rotation.detach(), "_rotation", device),
            _f32c(_scaling.detach(), "_scaling", device), _f32c(_opacity.detach(), "_opacity", device)]
    a.means3D, a.rotations, a.scales, a.opacities = (t.data_ptr() for t in keep)
    if binding is not None:
        keep += [binding.to(torch.int32).contiguous(), _f32c(face_center.detach(), "face_center", device),
                 _f32c(face_orien_mat.detach(), "face_orien_mat", device),
                 _f32c(face_scaling.detach(), "face_scaling", device)]
        a.binding, a.face_center, a.face_orien_mat, a.face_scaling = (t.data_ptr() for t in keep[4:])
        a.num_faces = keep[5].shape[0]
    e = lambda *s: torch.empty(s, dtype=torch.float32, device=device)  # noqa: E731
    means3D, opac, scales, cov = e(P, 3), e(P, 1), e(P, 3), e(P, 6)
    with torch.cuda.device(device):
        stream = torch.cuda.current_stream(device).cuda_stream
        N.check(N.lib().gab200_bind_activate(C.byref(a), means3D.data_ptr(), opac.data_ptr(), scales.data_ptr(),
                                             cov.data_ptr(), C.c_void_p(stream)), "gab200_bind_activate")
    return means3D, opac, scales, cov


# ================================================================================================================
# Per-face frame: one launch instead of ~25 eager ones, differentiable w.r.t. the vertices
# ================================================================================================================
class _FaceFrame(torch.autograd.Function):
    @staticmethod
    def forward(ctx, verts, faces):
        device = verts.device
        if device.type != "cuda":
            raise RuntimeError("gaussianavatars_b200 has no CPU path: tensors must be CUDA tensors")
        v = _f32c(verts.reshape(-1, 3), "verts", device)
        f = faces if faces.dtype == torch.int32 else faces.to(torch.int32)
        f = f.contiguous()
        V, F = v.shape[0], f.shape[0]
        fc = torch.empty((F, 3), dtype=torch.float32, device=device)
        fR = torch.empty((F, 3, 3), dtype=torch.float32, device=device)
        fs = torch.empty((F, 1), dtype=torch.float32, device=device)
        with torch.cuda.device(device):
            stream = torch.cuda.current_stream(device).cuda_stream
            N.check(N.lib().gab200_face_frame_forward(F, V, v.data_ptr(), f.data_ptr(), fc.data_ptr(), fR.data_ptr(),
                                                      fs.data_ptr(), C.c_void_p(stream)), "gab200_face_frame_forward")
        ctx.keep = (v, f, verts.shape)
        return fc, fR, fs

    @staticmethod
    def backward(ctx, g_fc, g_fR, g_fs):
        v, f, shape = ctx.keep
        device = v.device
        gv = torch.empty_like(v)
        c = lambda t: None if t is None else (t if t.is_contiguous() else t.contiguous())  # noqa: E731
        g_fc, g_fR, g_fs = c(g_fc), c(g_fR), c(g_fs)
        with torch.cuda.device(device):
            stream = torch.cuda.current_stream(device).cuda_stream
            N.check(N.lib().gab200_face_frame_backward(f.shape[0], v.shape[0], v.data_ptr(), f.data_ptr(), N.ptr(g_fc),
                                                       N.ptr(g_fR), N.ptr(g_fs), gv.data_ptr(), C.c_void_p(stream)),
                    "gab200_face_frame_backward")
        return gv.view(shape), None


def face_frame(verts: torch.Tensor, faces: torch.Tensor):
    """verts (V,3) [or (1,V,3)], faces (F,3) -> face_center (F,3), face_orien_mat (F,3,3), face_scaling (F,1).
    Replaces update_mesh_properties / compute_face_orientation (scene/flame_gaussian_model.py:137-147)."""
    return _FaceFrame.apply(verts, faces)


# ================================================================================================================
# L1 loss against a uint8 ground truth: loss and dL/dimage in one kernel
# ================================================================================================================
class _L1LossU8(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, gt_u8):
        device = image.device
        if device.type != "cuda":
            raise RuntimeError("gaussianavatars_b200 has no CPU path: tensors must be CUDA tensors")
        img = _f32c(image, "image", device)
        if gt_u8.dtype != torch.uint8 or gt_u8.numel() != img.numel():
            raise TypeError("gt must be a uint8 tensor with the image's number of elements")
        gt = gt_u8 if gt_u8.is_contiguous() else gt_u8.contiguous()
        loss = torch.empty((), dtype=torch.float32, device=device)
        with torch.cuda.device(device):
            stream = torch.cuda.current_stream(device).cuda_stream
            N.check(N.lib().gab200_l1_loss_u8(img.numel(), img.data_ptr(), gt.data_ptr(), None, loss.data_ptr(),
                                              C.c_void_p(stream)), "gab200_l1_loss_u8")
        ctx.save_for_backward(img, gt)
        return loss

    @staticmethod
    def backward(ctx, g):
        img, gt = ctx.saved_tensors
        device = img.device
        grad = torch.empty_like(img)
        gs = g if (g.dtype == torch.float32 and g.device == device) else g.to(device=device, dtype=torch.float32)
        with torch.cuda.device(device):
            stream = torch.cuda.current_stream(device).cuda_stream
            N.check(N.lib().gab200_l1_loss_u8_backward(img.numel(), img.data_ptr(), gt.data_ptr(), gs.data_ptr(),
                                                       grad.data_ptr(), C.c_void_p(stream)), "gab200_l1_loss_u8_backward")
        return grad, None


def l1_loss_u8(image: torch.Tensor, gt_u8: torch.Tensor) -> torch.Tensor:
    """mean |image - gt/255| for a uint8 ground truth (same shape), differentiable w.r.t. `image`."""
    return _L1LossU8.apply(image, gt_u8)
