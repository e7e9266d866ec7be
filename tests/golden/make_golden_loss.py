"""Generates tests/golden/loss_vectors.npz by IMPORTING the real reference loss (utils/loss_utils.py under
/root/reference, read-only) and by running torch.optim.Adam exactly as the reference configures it
(scene/gaussian_model.py:213-232: per-array parameter groups, lr per group, eps=1e-15).  The fixture travels to the
GPU box; /root/reference does not.

    python tests/golden/make_golden_loss.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
from utils.loss_utils import l1_loss, ssim  # noqa: E402  (REAL reference code)

LAMBDA = 0.2  # arguments/__init__.py:99 lambda_dssim


def image_pair(seed, C, H, W):
    """A render-like image (smooth blobs + mild noise, some flat background) and a uint8 ground truth near it."""
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.linspace(0, 1, H), torch.linspace(0, 1, W), indexing="ij")
    img = torch.zeros(C, H, W)
    for _ in range(6):
        cx, cy, r = torch.rand(3, generator=g)
        col = torch.rand(C, generator=g)
        blob = torch.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (0.02 + 0.1 * r) ** 2)
        img += col[:, None, None] * blob
    img = img.clamp(0, 1)
    gt = (img + 0.08 * torch.randn(C, H, W, generator=g)).clamp(0, 1)
    gt[:, : H // 4, : W // 3] = 1.0          # flat white background patch (sigma ~ 0: the stiff part of SSIM)
    img[:, : H // 4, : W // 5] = 1.0
    gt_u8 = (gt * 255).round().to(torch.uint8)
    return img.contiguous(), gt_u8.contiguous()


def run_loss(img, gt_u8, dtype):
    x = img.to(dtype).clone().requires_grad_(True)
    y = gt_u8.to(dtype) / 255
    l1 = l1_loss(x, y)
    s = ssim(x, y)
    total = l1 * (1.0 - LAMBDA) + (1.0 - s) * LAMBDA      # train.py:131-132
    total.backward()
    return float(l1), float(s), float(total), x.grad.detach().numpy()


def main():
    out = {}
    for name, (seed, C, H, W) in {"a": (0, 3, 45, 70), "b": (1, 3, 64, 96), "c": (2, 1, 33, 31)}.items():
        img, gt_u8 = image_pair(seed, C, H, W)
        out[f"{name}_image"] = img.numpy()
        out[f"{name}_gt_u8"] = gt_u8.numpy()
        for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
            l1, s, total, grad = run_loss(img, gt_u8, dt)
            out[f"{name}_{tag}_scalars"] = np.array([l1, s, total], dtype=np.float64)
            out[f"{name}_{tag}_grad"] = grad

    # ---- Adam: three groups with the reference's learning rates, 4 steps, fresh gradients each step
    g = torch.Generator().manual_seed(7)
    shapes = {"xyz": ((203, 3), 0.005), "f_rest": ((203, 15, 3), 0.0025 / 20.0), "opacity": ((203, 1), 0.05)}
    params = {k: torch.randn(*sh, generator=g).requires_grad_(True) for k, (sh, _) in shapes.items()}
    opt = torch.optim.Adam([{"params": [params[k]], "lr": lr, "name": k} for k, (_, lr) in shapes.items()],
                           lr=0.0, eps=1e-15)
    for k in shapes:
        out[f"adam_{k}_p0"] = params[k].detach().numpy().copy()
        out[f"adam_{k}_lr"] = np.array(shapes[k][1], dtype=np.float64)
    for step in range(1, 5):
        for k in shapes:
            grad = torch.randn(params[k].shape, generator=g) * (10.0 ** float(torch.randint(-6, 1, (1,), generator=g)))
            if step == 2:
                grad[::3] = 0.0                       # invisible splats receive exact zeros
            params[k].grad = grad
            out[f"adam_{k}_g{step}"] = grad.numpy().copy()
        opt.step()
        for k in shapes:
            out[f"adam_{k}_p{step}"] = params[k].detach().numpy().copy()
    for k in shapes:
        st = opt.state[params[k]]
        out[f"adam_{k}_m"] = st["exp_avg"].numpy().copy()
        out[f"adam_{k}_v"] = st["exp_avg_sq"].numpy().copy()
    path = os.path.join(HERE, "loss_vectors.npz")
    np.savez_compressed(path, **out)
    print(path, len(out), "arrays", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
