"""GPU: the per-splat depth sort has two implementations behind gab200_forward -- cub radix sort + scan (no hint) and
a bucket sort over a hinted key range.  Both must produce the reference's order bit for bit (ties by ascending splat
id), whatever the hint says."""
import numpy as np
import pytest
import torch

from tests import helpers as h
from tests.test_gpu_parity import _run_cuda, _dev

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["counting-tile-sort", "cub-radix-tile-sort"])
def tile_sort_implementation(request):
    from gaussianavatars_b200 import _native as N

    N.tune(N.TUNE_TILE_SORT, 1 if request.param.startswith("counting") else 0)
    yield
    N.tune(N.TUNE_TILE_SORT, 0)


def _stream(scene, dev, hint=None, exact=True):
    """One forward; returns (image, radii, keys, vals, ranges, n, depth_sort_path)."""
    from gaussianavatars_b200 import rasterizer as R

    key = (dev, int(scene["W"]), int(scene["H"]), int(scene["means3D"].shape[0]))
    if hint == "none":
        R._default_hints.set_depth(key, (0, 0))
    elif hint is not None:
        R._default_hints.set_depth(key, hint)
    img, radii, _, _ = _run_cuda(scene, dev, exact=exact)
    keys, vals, ranges, n = R.export_last_binning()
    path = int(R._last[1].depth_sort_path)
    return (img.cpu().numpy(), radii.cpu().numpy(), keys.cpu().numpy().view(np.uint64), vals.cpu().numpy().view(np.uint32),
            ranges.cpu().numpy().view(np.uint32), n, path)


def _same(a, b):
    for x, y, what in zip(a[:5], b[:5], ("image", "radii", "keys", "values", "ranges")):
        assert np.array_equal(x, y), f"{what} differ between the radix and the bucket depth sort"
    assert a[5] == b[5]


@pytest.mark.parametrize("P,W,H,deg,seed,exact", [(10_000, 256, 256, 0, 0, True), (4_000, 333, 201, 3, 1, False),
                                                   (50_000, 640, 360, 0, 5, False), (300, 64, 48, 1, 2, True)])
def test_bucket_depth_sort_equals_radix_sort_and_the_oracle(P, W, H, deg, seed, exact):
    dev = _dev()
    scene = h.random_scene(P, W, H, sh_degree=deg, seed=seed, scale_shift=0.6)
    first = _stream(scene, dev, hint="none", exact=exact)        # no hint -> radix path, leaves a hint behind
    assert first[6] == 0
    second = _stream(scene, dev, exact=exact)                      # hinted -> bucket path
    assert second[6] == 1, "the hinted frame did not take the bucket path"
    _same(first, second)
    if exact:
        st = h.oracle_forward(scene)
        assert second[5] == st.N
        assert np.array_equal(second[2], st.keys_sorted) and np.array_equal(second[3], st.vals_sorted)
        assert np.array_equal(second[4], st.ranges)


def test_wrong_hints_never_change_the_result():
    dev = _dev()
    scene = h.random_scene(20_000, 320, 200, sh_degree=0, seed=7, scale_shift=0.4)
    ref = _stream(scene, dev, hint="none")
    lo, hi = ref[2].astype(np.uint64) & np.uint64(0xFFFFFFFF), None
    kmin, kmax = int(lo.min()), int(lo.max())
    mid = (kmin + kmax) // 2
    paths = []
    for hint in [(1, 2),                               # everything above the range: one end bucket -> overflow -> redo
                 (0xFFFFFFF0, 0xFFFFFFFE),             # everything below the range
                 (1, 0xFFFFFFFE),                      # absurdly wide: a handful of buckets hold everything
                 (kmin, mid),                          # upper half clamped into the last bucket
                 (mid, kmax + 1),                      # lower half clamped into the first bucket
                 (kmin, kmax)]:                        # tight
        out = _stream(scene, dev, hint=hint)
        _same(ref, out)
        paths.append(out[6])
    assert paths[0] == 2 and paths[1] == 2, paths     # 20k splats cannot fit one 2048-splat bucket
    assert paths[-1] == 1, paths


def test_bucket_depth_sort_breaks_ties_by_splat_id():
    """Thousands of splats at exactly equal depths (several per bucket, and one depth shared by > 32 splats)."""
    dev = _dev()
    scene = h.random_scene(6_000, 200, 120, sh_degree=0, seed=9, scale_shift=0.5)
    z = scene["means3D"][:, 2]
    levels = torch.linspace(float(z.min()), float(z.max()), 40)
    scene["means3D"][:, 2] = levels[torch.randint(0, 40, (z.numel(),), generator=torch.Generator().manual_seed(1))]
    st = h.oracle_forward(scene)
    a = _stream(scene, dev, hint="none")
    b = _stream(scene, dev)
    assert b[6] == 1
    _same(a, b)
    assert np.array_equal(b[2], st.keys_sorted) and np.array_equal(b[3], st.vals_sorted)


def test_bucket_path_with_nothing_visible_and_with_backward():
    import gaussianavatars_b200 as g
    from gaussianavatars_b200 import rasterizer as R

    dev = _dev()
    scene = h.random_scene(2_000, 96, 64, sh_degree=1, seed=11, scale_shift=0.5)
    ref = _stream(scene, dev, hint="none")
    # gradients through a bucket-sorted frame equal those through a radix-sorted one (same sorted stream -> same math)
    grads = []
    for hint in ("none", None):
        if hint == "none":
            R._default_hints.shapes.clear()
        img, radii, t, m2 = _run_cuda(scene, dev, need_grad=True)
        (img * torch.linspace(0, 1, img.numel(), device=dev).view_as(img)).sum().backward()
        grads.append([t[k].grad.clone() for k in ("means3D", "scales", "rotations", "opacities", "shs")] + [m2.grad.clone()])
    assert int(R._last[1].depth_sort_path) == 1
    for ga, gb in zip(*grads):
        assert torch.allclose(ga, gb, rtol=1e-4, atol=1e-7)
    # every splat behind the near plane, with a hint present: M = 0, N = 0, image = background
    behind = dict(scene)
    behind["means3D"] = scene["means3D"].clone()
    behind["means3D"][:, 2] = 0.1
    out = _stream(behind, dev)
    assert out[5] == 0 and int(np.abs(out[1]).sum()) == 0
    assert np.allclose(out[0], scene["bg"].numpy()[:, None, None])
    # and the next frame of the original scene is unaffected by whatever that left behind
    _same(ref, _stream(scene, dev))
