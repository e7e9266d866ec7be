"""Generates tests/golden/densify_vectors.npz by running the REAL reference densification on CPU tensors:
    scene/gaussian_model.py:334-519  (replace/prune/cat optimizer surgery, prune_points, densification_postfix,
                                       densify_and_split, densify_and_clone, densify_and_prune)
through tests/ref_import.py.  The reference hard-codes device="cuda" in its tensor factories and draws the split
samples with torch.normal; for the duration of the call
  * torch.zeros / torch.ones_like / torch.zeros_like lose a device="cuda" keyword (nothing else changes), and
  * torch.normal(mean=0, std=stds) is replaced by `mean + noise * std` with `noise` a recorded standard-normal tensor
    (what torch.normal computes internally), so that the CUDA op can be fed the SAME noise.
Two cases: a mesh-bound model (binding / binding_counter / face_scaling set as FlameGaussianModel does) with the
screen-size criterion on, and an unbound model with max_screen_size=None.

    python tests/golden/make_golden_densify.py
"""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)

from tests import ref_import  # noqa: E402

NAMES = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")
ATTR = {"xyz": "_xyz", "f_dc": "_features_dc", "f_rest": "_features_rest", "opacity": "_opacity",
        "scaling": "_scaling", "rotation": "_rotation"}


class _Patches:
    """device="cuda" -> CPU for the factories the densification code uses; torch.normal with recorded noise."""

    def __init__(self, gen):
        self.gen, self.noise = gen, []
        self.saved = {}

    def __enter__(self):
        for name in ("zeros", "ones_like", "zeros_like", "ones"):
            real = getattr(torch, name)
            self.saved[name] = real

            def wrap(*a, _real=real, **k):
                if k.get("device") == "cuda":
                    k.pop("device")
                return _real(*a, **k)

            setattr(torch, name, wrap)
        self.saved["normal"] = torch.normal

        def normal(mean=None, std=None, **k):
            n = torch.randn(std.shape, generator=self.gen, dtype=std.dtype)
            self.noise.append(n)
            return mean + n * std

        torch.normal = normal
        self.saved["empty_cache"] = torch.cuda.empty_cache
        torch.cuda.empty_cache = lambda: None
        return self

    def __exit__(self, *a):
        for name in ("zeros", "ones_like", "zeros_like", "ones", "normal"):
            setattr(torch, name, self.saved[name])
        torch.cuda.empty_cache = self.saved["empty_cache"]


def build_model(P, F, sh_degree, bound, seed):
    from scene.gaussian_model import GaussianModel

    g = torch.Generator().manual_seed(seed)
    M = (sh_degree + 1) ** 2
    m = GaussianModel(sh_degree)
    nn = torch.nn
    m._xyz = nn.Parameter(torch.randn(P, 3, generator=g) * 0.5)
    m._features_dc = nn.Parameter(torch.randn(P, 1, 3, generator=g))
    m._features_rest = nn.Parameter(torch.randn(P, M - 1, 3, generator=g) * 0.1)
    m._opacity = nn.Parameter(torch.randn(P, 1, generator=g) * 3.0 - 2.0)       # plenty below sigmoid^-1(0.005) = -5.3
    m._scaling = nn.Parameter(torch.randn(P, 3, generator=g) * 1.2 - 1.0)
    m._rotation = nn.Parameter(torch.randn(P, 4, generator=g))
    m.max_radii2D = torch.rand(P, generator=g) * 40.0
    if bound:
        binding = torch.randint(0, F, (P,), generator=g)
        binding[: F // 2] = torch.arange(F // 2)           # the other half of the faces is sparsely populated
        m.binding = binding
        m.binding_counter = torch.bincount(binding, minlength=F).to(torch.int32)
        m.face_scaling = torch.rand(F, 1, generator=g) * 0.02 + 0.002
        # the rest of the per-face frame (get_xyz is only asked for its row count here): identity frames
        m.face_center = torch.zeros(F, 3)
        m.face_orien_mat = torch.eye(3).repeat(F, 1, 1)
        m.face_orien_quat = torch.tensor([1.0, 0.0, 0.0, 0.0]).repeat(F, 1)
    args = SimpleNamespace(percent_dense=0.01, position_lr_init=0.005, position_lr_final=0.00005, position_lr_delay_mult=0.01,
                           position_lr_max_steps=30000, feature_lr=0.0025, opacity_lr=0.05, scaling_lr=0.017, rotation_lr=0.001)
    m.spatial_lr_scale = 1.0
    ref_import.on_cpu(m.training_setup, args)
    # a few optimizer steps so that exp_avg / exp_avg_sq are populated
    for it in range(3):
        for n in NAMES:
            p = getattr(m, ATTR[n])
            p.grad = torch.randn(p.shape, generator=g) * 0.01
        m.optimizer.step()
        m.optimizer.zero_grad(set_to_none=True)
    m.xyz_gradient_accum = torch.rand(P, 1, generator=g) * 0.001
    m.denom = torch.randint(0, 4, (P, 1), generator=g).float()   # zeros included: 0/0 -> nan -> 0
    return m


def snapshot(m, prefix, out):
    for n in NAMES:
        p = getattr(m, ATTR[n])
        out[f"{prefix}_{n}"] = p.detach().numpy().copy()
        st = m.optimizer.state.get(p)
        out[f"{prefix}_{n}_exp_avg"] = st["exp_avg"].numpy().copy()
        out[f"{prefix}_{n}_exp_avg_sq"] = st["exp_avg_sq"].numpy().copy()
    out[f"{prefix}_xyz_gradient_accum"] = m.xyz_gradient_accum.numpy().copy()
    out[f"{prefix}_denom"] = m.denom.numpy().copy()
    out[f"{prefix}_max_radii2D"] = m.max_radii2D.numpy().copy()
    if getattr(m, "binding", None) is not None:
        out[f"{prefix}_binding"] = m.binding.numpy().astype(np.int32)
        out[f"{prefix}_binding_counter"] = m.binding_counter.numpy().astype(np.int32)
        out[f"{prefix}_face_scaling"] = m.face_scaling.numpy().copy()


def main():
    ref_import.prepare()
    out = {}
    for case, (P, F, deg, bound, max_grad, min_op, extent, screen, seed) in {
            "bound": (1500, 120, 1, True, 0.0002, 0.005, 0.6, 20, 11),
            "plain": (900, 0, 0, False, 0.0002, 0.005, 4.0, None, 12)}.items():
        m = build_model(P, F, deg, bound, seed)
        snapshot(m, f"{case}_in", out)
        gen = torch.Generator().manual_seed(100 + seed)
        with _Patches(gen) as pt:
            m.densify_and_prune(max_grad, min_op, extent, screen)
        snapshot(m, f"{case}_out", out)
        out[f"{case}_noise"] = (torch.cat(pt.noise) if pt.noise else torch.zeros(0, 3)).numpy()
        out[f"{case}_hyper"] = np.array([max_grad, min_op, extent, -1.0 if screen is None else float(screen), 0.01], np.float64)
        print(case, "P", P, "->", m._xyz.shape[0], "split parents", out[f"{case}_noise"].shape[0] // 2)
    path = os.path.join(HERE, "densify_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
