"""The two callers either side of the rasterizer in the reference's training step, as
CUDA launches behind the same C ABI:

  photometric_loss(image, gt, lambda_dssim)   <- l1_loss * (1 - lambda) + (1 - ssim) * lambda
                                                 (utils/loss_utils.py:17-18,36-63, train.py:131-132)
  Adam(param_groups, lr, betas, eps)          <- torch.optim.Adam(l, lr=0.0, eps=1e-15).step()
                                                 (scene/gaussian_model.py:213-232, train.py:207-210)

No CPU or eager fallback: both raise when the library is missing or a tensor is not a CUDA float32 tensor.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _native as N


# ================================================================================================================
# (1 - lambda) L1 + lambda (1 - SSIM), loss and dL/dimage in two launches
# ================================================================================================================
class _PhotometricLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, gt, lambda_dssim):
        device = image.device
        if device.type != "cuda":
            raise RuntimeError("gaussianavatars_b200 has no CPU path: tensors must be CUDA tensors")
        if image.dtype != torch.float32 or image.dim() not in (3, 4):
            raise TypeError("image must be a float32 (C, H, W) or (B, C, H, W) tensor")
        if gt.shape != image.shape or gt.device != device:
            raise ValueError(f"gt must have the image's shape {tuple(image.shape)} on {device}, got {tuple(gt.shape)} on {gt.device}")
        if gt.dtype not in (torch.uint8, torch.float32):
            raise TypeError(f"gt must be uint8 (value/255) or float32, got {gt.dtype}")
        if not 0.0 <= float(lambda_dssim) <= 1.0:
            raise ValueError("lambda_dssim must lie in [0, 1]")
        img = image if image.is_contiguous() else image.contiguous()
        g = gt if gt.is_contiguous() else gt.contiguous()
        H, W = int(img.shape[-2]), int(img.shape[-1])
        Cc = img.numel() // max(H * W, 1)   # a batch is just more independent planes: the means run over all of them
        grad = torch.empty_like(img)
        scratch = torch.empty(N.PHOTOMETRIC_SCRATCH_HEAD + 3 * img.numel(), dtype=torch.float32, device=device)
        loss = torch.empty(3, dtype=torch.float32, device=device)
        a = N.PhotometricArgs()
        a.abi_version = N.ABI_VERSION
        a.channels, a.height, a.width = Cc, H, W
        a.gt_is_u8 = 1 if gt.dtype == torch.uint8 else 0
        a.lambda_dssim = float(lambda_dssim)
        a.image, a.gt, a.grad = img.data_ptr(), g.data_ptr(), grad.data_ptr()
        a.loss, a.scratch = loss.data_ptr(), scratch.data_ptr()
        with torch.cuda.device(device):
            stream = torch.cuda.current_stream(device).cuda_stream
            N.check(N.lib().gab200_photometric_loss(C.byref(a), C.c_void_p(stream)), "gab200_photometric_loss")
        ctx.save_for_backward(grad)
        ctx.mark_non_differentiable(loss)
        return loss[2], loss

    @staticmethod
    def backward(ctx, g_total, _g_parts):
        (grad,) = ctx.saved_tensors
        return grad.mul_(g_total), None, None  # in place: the buffer is ours and single-use


def photometric_loss(image: torch.Tensor, gt: torch.Tensor, lambda_dssim: float = 0.2, return_parts: bool = False):
    """`(1 - lambda_dssim) * l1_loss(image, gt) + lambda_dssim * (1 - ssim(image, gt))` of the reference training
    loop, differentiable w.r.t. `image` (float32, (C, H, W) or (B, C, H, W) like the reference's `ssim`).  `gt` is float32 in [0, 1] like `viewpoint_cam.original_image`
    or the uint8 image it was decoded from (value/255 in-kernel: a quarter of the upload).  With `return_parts` also
    returns the detached tensor [l1 mean, ssim mean, total] (for logging, train.py:159-166)."""
    total, parts = _PhotometricLoss.apply(image, gt, lambda_dssim)
    return (total, parts) if return_parts else total


# ================================================================================================================
# Adam: all parameter groups in one launch
# ================================================================================================================
class Adam(torch.optim.Optimizer):
    """Drop-in for `torch.optim.Adam(param_groups, lr=0.0, eps=1e-15)` as the reference builds it.

    Same `param_groups` / `state` layout as torch's Adam (`state[p] = {"step", "exp_avg", "exp_avg_sq"}`), so the
    reference's densification surgery on the optimizer state (scene/gaussian_model.py:334-419) and
    `optimizer.state_dict()` checkpoints (scene/gaussian_model.py:89,111) work unchanged.  amsgrad, weight decay and
    maximize are not part of the reference's configuration and are rejected."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, maximize=False):
        if weight_decay != 0 or amsgrad or maximize:
            raise ValueError("gaussianavatars_b200.Adam implements the reference configuration only "
                             "(weight_decay=0, amsgrad=False, maximize=False)")
        if not 0.0 <= lr or not 0.0 <= eps or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        batches = {}   # (device, step, beta1, beta2, eps) -> list of segments; one launch per 8 segments
        keep = []
        for group in self.param_groups:
            beta1, beta2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.device.type != "cuda" or p.dtype != torch.float32 or p.grad.is_sparse or not p.is_contiguous():
                    raise RuntimeError("gaussianavatars_b200.Adam steps contiguous CUDA float32 parameters only "
                                       "(no CPU or eager fallback)")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                m, v = st["exp_avg"], st["exp_avg_sq"]
                if not (m.is_contiguous() and v.is_contiguous()):
                    raise RuntimeError("Adam state tensors must be contiguous")
                keep.append(g)
                seg = N.AdamSegment(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), float(group["lr"]))
                key = (p.device, int(st["step"]), float(beta1), float(beta2), float(group["eps"]))
                batches.setdefault(key, []).append(seg)
        for (device, step, beta1, beta2, eps), segs in batches.items():
            arr = (N.AdamSegment * len(segs))(*segs)
            with torch.cuda.device(device):
                stream = torch.cuda.current_stream(device).cuda_stream
                N.check(N.lib().gab200_adam_step(len(segs), arr, step, beta1, beta2, eps, C.c_void_p(stream)),
                        "gab200_adam_step")
        return loss


# ================================================================================================================
# position / scale regularisers of the mesh-bound training step (train.py:134-146), loss + gradient in one launch each
# ================================================================================================================
class _BindingRegularizers(torch.autograd.Function):
    @staticmethod
    def forward(ctx, _xyz, _scaling, face_scaling, radii, binding, thr_xyz, thr_scale, lam_xyz, lam_scale, metric_xyz,
                metric_scale):
        device = _xyz.device
        if device.type != "cuda":
            raise RuntimeError("gaussianavatars_b200 has no CPU path: tensors must be CUDA tensors")
        for t, n in ((_xyz, "_xyz"), (_scaling, "_scaling")):
            if t.dtype != torch.float32 or t.shape != (_xyz.shape[0], 3):
                raise TypeError(f"{n} must be a float32 (P, 3) tensor")
        P = _xyz.shape[0]
        x = _xyz if _xyz.is_contiguous() else _xyz.contiguous()
        s = _scaling if _scaling.is_contiguous() else _scaling.contiguous()
        r = radii if radii.dtype == torch.int32 else radii.to(torch.int32)   # a bool visibility_filter works as well
        r = r if r.is_contiguous() else r.contiguous()
        if r.numel() != P:
            raise ValueError("radii / visibility_filter must have one entry per splat")
        a = N.RegularizeArgs()
        a.abi_version, a.P = N.ABI_VERSION, P
        a.metric_xyz, a.metric_scale = int(bool(metric_xyz)), int(bool(metric_scale))
        a.threshold_xyz, a.threshold_scale = float(thr_xyz), float(thr_scale)
        a.lambda_xyz, a.lambda_scale = float(lam_xyz), float(lam_scale)
        a.xyz, a.scaling, a.radii = x.data_ptr(), s.data_ptr(), r.data_ptr()
        keep = [x, s, r]
        if binding is not None:
            b = binding if binding.dtype == torch.int32 else binding.to(torch.int32)
            fs = face_scaling.reshape(-1)
            fs = fs if fs.is_contiguous() else fs.contiguous()
            a.binding, a.face_scaling = b.data_ptr(), fs.data_ptr()
            keep += [b, fs]
        loss = torch.empty(3, dtype=torch.float32, device=device)
        sums = torch.empty(3, dtype=torch.float64, device=device)
        a.loss, a.sums = loss.data_ptr(), sums.data_ptr()
        with torch.cuda.device(device):
            stream = torch.cuda.current_stream(device).cuda_stream
            N.check(N.lib().gab200_regularize_forward(C.byref(a), C.c_void_p(stream)), "gab200_regularize_forward")
        ctx.args, ctx.keep, ctx.sums = a, keep, sums
        ctx.face_shape = None if face_scaling is None else face_scaling.shape
        ctx.want_face = binding is not None and face_scaling is not None and ctx.needs_input_grad[2] and \
            (metric_xyz or metric_scale)
        ctx.mark_non_differentiable(loss)
        return loss[0], loss[1], loss

    @staticmethod
    def backward(ctx, g_xyz, g_scale, _g_all):
        a = ctx.args
        device = ctx.keep[0].device
        P = a.P
        gx = torch.empty((P, 3), dtype=torch.float32, device=device)
        gs = torch.empty((P, 3), dtype=torch.float32, device=device)
        gf = torch.zeros(ctx.face_shape, dtype=torch.float32, device=device) if ctx.want_face else None
        z = torch.zeros((), dtype=torch.float32, device=device)
        g = torch.stack((g_xyz if g_xyz is not None else z, g_scale if g_scale is not None else z)).float().contiguous()
        a.grad_xyz, a.grad_scaling, a.grad_face_scaling = gx.data_ptr(), gs.data_ptr(), N.ptr(gf)
        with torch.cuda.device(device):
            stream = torch.cuda.current_stream(device).cuda_stream
            N.check(N.lib().gab200_regularize_backward(C.byref(a), g.data_ptr(), C.c_void_p(stream)),
                    "gab200_regularize_backward")
        return gx, gs, gf, None, None, None, None, None, None, None, None


def binding_regularizers(_xyz, _scaling, radii, binding=None, face_scaling=None, threshold_xyz=1.0, threshold_scale=0.6,
                         lambda_xyz=1e-2, lambda_scale=1.0, metric_xyz=False, metric_scale=False, return_count=False):
    """`losses['xyz']`, `losses['scale']` of the reference training step (train.py:134-146; defaults
    arguments/__init__.py:100-105), differentiable w.r.t. `_xyz`, `_scaling` (and `face_scaling` in the metric
    variants).  `radii` is the rendered frame's radii (or its `visibility_filter`).  A training step that adds these
    two to the photometric loss never calls `get_scaling` / boolean-mask indexing."""
    lx, ls, all_ = _BindingRegularizers.apply(_xyz, _scaling, face_scaling, radii, binding, threshold_xyz, threshold_scale,
                                              lambda_xyz, lambda_scale, metric_xyz, metric_scale)
    return (lx, ls, all_[2]) if return_count else (lx, ls)
