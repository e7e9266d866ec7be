"""GPU parity of the two training-step neighbours of the rasterizer, through the C ABI:
photometric loss (L1 + D-SSIM, forward + gradient) and the multi-tensor Adam step.  Golden vectors come from the real
reference (tests/golden/make_golden_loss.py)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "loss_vectors.npz"))
LAMBDA = 0.2
SCALAR_TOL = 1e-6        # abs, on l1 / ssim / total (the reference's own fp32 run sits 7e-8 from its fp64 run)
GRAD_TOL = 1e-4          # x max|grad|  (reference fp32 vs fp64: 6e-6)


def _g():
    import gaussianavatars_b200 as g
    return g


@pytest.mark.parametrize("case", ["a", "b", "c"])
@pytest.mark.parametrize("gt_kind", ["u8", "f32"])
def test_photometric_loss_matches_reference_vectors(case, gt_kind):
    g = _g()
    img = torch.tensor(GOLD[f"{case}_image"], device="cuda").requires_grad_(True)
    gt = torch.tensor(GOLD[f"{case}_gt_u8"], device="cuda")
    if gt_kind == "f32":
        gt = (gt.cpu().float() / 255).cuda()
    total, parts = g.photometric_loss(img, gt, LAMBDA, return_parts=True)
    total.backward()
    ref = GOLD[f"{case}_f64_scalars"]
    assert np.abs(parts.cpu().numpy().astype(np.float64) - ref).max() < SCALAR_TOL
    assert abs(float(total.detach()) - ref[2]) < SCALAR_TOL
    gref = GOLD[f"{case}_f64_grad"]
    err = np.abs(img.grad.cpu().numpy().astype(np.float64) - gref).max()
    assert err < GRAD_TOL * np.abs(gref).max(), err / np.abs(gref).max()


@pytest.mark.parametrize("shape", [(3, 1080, 1920), (3, 802, 550), (3, 37, 1), (1, 1, 45), (3, 32, 32), (3, 33, 65)])
def test_photometric_loss_matches_torch_restatement(shape):
    """Full BASELINE sizes and ragged edges against the conv2d restatement of the reference (fp32 and fp64 on the GPU)."""
    from oracle import loss as ol
    g = _g()
    gen = torch.Generator(device="cuda").manual_seed(3)
    C, H, W = shape
    base = torch.rand(C, (H + 7) // 8, (W + 7) // 8, device="cuda", generator=gen)
    img = torch.nn.functional.interpolate(base[None], size=(H, W), mode="bilinear", align_corners=False)[0].contiguous()
    gt_u8 = ((img + 0.05 * torch.randn(C, H, W, device="cuda", generator=gen)).clamp(0, 1) * 255).round().to(torch.uint8)
    gt_u8[:, : H // 3, : W // 3] = 255
    img = img.clone()
    img[:, : H // 3, : W // 4] = 1.0
    x = img.clone().requires_grad_(True)
    total, parts = g.photometric_loss(x, gt_u8, LAMBDA, return_parts=True)
    total.backward()
    xd = img.double().requires_grad_(True)
    l1, ssim, tot = ol.photometric_torch(xd, (gt_u8.cpu().float() / 255).cuda().double(), LAMBDA)   # fp32 IEEE division on the CPU, as the
    # reference's loader does (utils/general_utils.py:21-23); torch's CUDA `/ 255` multiplies by a reciprocal instead
    tot.backward()
    ref = torch.stack([l1, ssim, tot]).detach()
    assert (parts.double() - ref).abs().max().item() < 2e-6
    gmax = xd.grad.abs().max().item()
    err = (x.grad.double() - xd.grad).abs().max().item()
    assert err < GRAD_TOL * gmax, (err / gmax)


def test_photometric_loss_properties():
    g = _g()
    gen = torch.Generator(device="cuda").manual_seed(5)
    img = torch.rand(3, 270, 480, device="cuda", generator=gen)
    gt_u8 = (torch.rand(3, 270, 480, device="cuda", generator=gen) * 255).to(torch.uint8)
    # lambda = 0 is the L1 kernel
    x = img.clone().requires_grad_(True)
    t0, p0 = g.photometric_loss(x, gt_u8, 0.0, return_parts=True)
    t0.backward()
    x1 = img.clone().requires_grad_(True)
    l1 = g.l1_loss_u8(x1, gt_u8)
    l1.backward()
    assert abs(float(t0.detach()) - float(l1.detach())) < 1e-6
    assert torch.equal(x.grad, x1.grad)
    # identical images: SSIM = 1, loss = 0, gradient ~ 0
    q = (gt_u8.cpu().float() / 255).cuda().requires_grad_(True)
    t, p = g.photometric_loss(q, gt_u8, LAMBDA, return_parts=True)
    t.backward()
    assert abs(float(p[1]) - 1.0) < 1e-6 and float(p[0]) == 0.0 and abs(float(t.detach())) < 1e-6
    assert q.grad.abs().max().item() < 1e-4 / q.numel() * 50
    # the total is linear in lambda
    ts = [float(g.photometric_loss(img, gt_u8, lam).detach()) for lam in (0.0, 0.5, 1.0)]
    assert abs(ts[1] - 0.5 * (ts[0] + ts[2])) < 1e-6
    # upstream gradient scaling
    x2 = img.clone().requires_grad_(True)
    (3.0 * g.photometric_loss(x2, gt_u8, LAMBDA)).backward()
    x3 = img.clone().requires_grad_(True)
    g.photometric_loss(x3, gt_u8, LAMBDA).backward()
    assert torch.allclose(x2.grad, 3.0 * x3.grad, rtol=1e-6, atol=0)


def test_photometric_loss_argument_errors():
    g = _g()
    img = torch.rand(3, 8, 8, device="cuda")
    with pytest.raises(ValueError):
        g.photometric_loss(img, torch.zeros(3, 8, 9, device="cuda", dtype=torch.uint8))
    with pytest.raises(TypeError):
        g.photometric_loss(img, torch.zeros(3, 8, 8, device="cuda", dtype=torch.int32))
    with pytest.raises(ValueError):
        g.photometric_loss(img, torch.zeros(3, 8, 8, device="cuda", dtype=torch.uint8), 1.5)
    with pytest.raises(RuntimeError):
        g.photometric_loss(img.cpu(), torch.zeros(3, 8, 8, dtype=torch.uint8))


NAMES = ("xyz", "f_rest", "opacity")


def test_adam_matches_reference_vectors():
    g = _g()
    params = {k: torch.nn.Parameter(torch.tensor(GOLD[f"adam_{k}_p0"], device="cuda")) for k in NAMES}
    opt = g.Adam([{"params": [params[k]], "lr": float(GOLD[f"adam_{k}_lr"]), "name": k} for k in NAMES], lr=0.0, eps=1e-15)
    for step in range(1, 5):
        for k in NAMES:
            params[k].grad = torch.tensor(GOLD[f"adam_{k}_g{step}"], device="cuda")
        opt.step()
        for k in NAMES:
            ref = GOLD[f"adam_{k}_p{step}"]
            err = np.abs(params[k].detach().cpu().numpy() - ref).max()
            assert err < 2e-6 * max(1.0, np.abs(ref).max()), (k, step, err)
    for k in NAMES:
        st = opt.state[params[k]]
        assert int(st["step"]) == 4
        assert np.allclose(st["exp_avg"].cpu().numpy(), GOLD[f"adam_{k}_m"], rtol=1e-5, atol=1e-12)
        assert np.allclose(st["exp_avg_sq"].cpu().numpy(), GOLD[f"adam_{k}_v"], rtol=1e-5, atol=1e-20)


def test_adam_matches_torch_adam_at_full_size_and_shares_its_state_layout():
    """100k splats x the six reference groups + three tiny FLAME-like groups (more than 8 segments: two launches),
    unaligned views, a parameter without gradient, state_dict round trip into torch.optim.Adam."""
    g = _g()
    gen = torch.Generator(device="cuda").manual_seed(11)
    P = 100_000
    shapes = [(P, 3), (P, 1, 3), (P, 15, 3), (P, 1), (P, 3), (P, 4), (7, 6), (1, 3), (13, 100), (5,)]
    lrs = [1.6e-4, 2.5e-3, 1.25e-4, 5e-2, 5e-3, 1e-3, 1e-3, 1e-6, 1e-3, 1e-2]
    init = [torch.randn(*s, device="cuda", generator=gen) for s in shapes]
    flat = torch.zeros(sum(t.numel() for t in init) + 1, device="cuda")

    def make(cls):
        ps = [torch.nn.Parameter(t.clone()) for t in init]
        return ps, cls([{"params": [p], "lr": lr, "name": str(i)} for i, (p, lr) in enumerate(zip(ps, lrs))], lr=0.0, eps=1e-15)

    ours, opt = make(g.Adam)
    theirs, ref = make(torch.optim.Adam)
    for step in range(3):
        off = 1                                          # gradient views at a 4-byte (not 16-byte) aligned offset
        flat.normal_(generator=gen)
        flat.mul_(10.0 ** (-step * 2))
        for i, (a, b) in enumerate(zip(ours, theirs)):
            if i == len(ours) - 1 and step == 0:
                a.grad = b.grad = None                   # skipped like torch does
                continue
            gview = flat[off:off + a.numel()].view_as(a)
            off += a.numel()
            a.grad = gview
            b.grad = gview.clone()
        opt.step()
        ref.step()
        for a, b in zip(ours, theirs):
            assert (a - b).abs().max().item() < 2e-6 * max(1.0, b.abs().max().item())
    for a, b in zip(ours, theirs):
        if a in opt.state:
            assert int(opt.state[a]["step"]) == int(ref.state[b]["step"])
            assert torch.allclose(opt.state[a]["exp_avg_sq"], ref.state[b]["exp_avg_sq"], rtol=1e-5, atol=1e-20)
    # our state_dict loads into torch's Adam and vice versa (checkpoints: scene/gaussian_model.py:89,111)
    ref.load_state_dict(opt.state_dict())
    opt.load_state_dict(ref.state_dict())
    for a in ours:
        a.grad = torch.ones_like(a)
    opt.step()


def test_adam_rejects_what_it_does_not_implement():
    g = _g()
    p = torch.nn.Parameter(torch.zeros(4, device="cuda"))
    with pytest.raises(ValueError):
        g.Adam([p], amsgrad=True)
    with pytest.raises(ValueError):
        g.Adam([p], weight_decay=0.1)
    q = torch.nn.Parameter(torch.zeros(4))
    q.grad = torch.ones(4)
    with pytest.raises(RuntimeError):
        g.Adam([q]).step()


def test_photometric_loss_accepts_a_batch_like_the_reference_ssim():
    """(B, C, H, W): every plane is convolved on its own and the means run over everything, so the batched total is
    the mean of the per-image totals and each gradient is 1/B of the per-image one (utils/loss_utils.py:36-63)."""
    g = _g()
    gen = torch.Generator(device="cuda").manual_seed(21)
    imgs = torch.rand(2, 3, 40, 70, device="cuda", generator=gen)
    gts = (torch.rand(2, 3, 40, 70, device="cuda", generator=gen) * 255).to(torch.uint8)
    x = imgs.clone().requires_grad_(True)
    total, parts = g.photometric_loss(x, gts, LAMBDA, return_parts=True)
    total.backward()
    singles, grads = [], []
    for b in range(2):
        xb = imgs[b].clone().requires_grad_(True)
        tb = g.photometric_loss(xb, gts[b], LAMBDA)
        tb.backward()
        singles.append(float(tb.detach()))
        grads.append(xb.grad)
    assert abs(float(total.detach()) - 0.5 * (singles[0] + singles[1])) < 1e-6
    assert x.grad.shape == imgs.shape
    assert torch.allclose(x.grad, 0.5 * torch.stack(grads), rtol=1e-5, atol=1e-12)


@pytest.mark.parametrize("metric_xyz,metric_scale,lam_scale", [(False, False, 1.0), (True, True, 1.0), (False, True, 0.5),
                                                              (True, False, 0.0)])
def test_binding_regularizers_match_the_training_step_lines(metric_xyz, metric_scale, lam_scale):
    """train.py:134-146 restated in eager float64 torch (the reference evaluates these lines inline in its loop)."""
    import gaussianavatars_b200 as g

    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(0)
    P, F = 50_000, 977
    xyz = (torch.randn(P, 3, generator=gen) * 0.9).to(dev).requires_grad_(True)
    sc = (torch.randn(P, 3, generator=gen) * 0.8 - 0.6).to(dev).requires_grad_(True)
    fs = (torch.rand(F, 1, generator=gen) * 2.0 + 0.2).to(dev).requires_grad_(True)
    binding = torch.randint(0, F, (P,), generator=gen).to(dev)
    radii = torch.randint(0, 3, (P,), generator=gen, dtype=torch.int32).to(dev)
    thr_x, thr_s, lam_x = 1.0, 0.6, 1e-2
    lx, ls, cnt = g.binding_regularizers(xyz, sc, radii, binding, fs, thr_x, thr_s, lam_x, lam_scale, metric_xyz, metric_scale,
                                         return_count=True)
    (3.0 * lx + 0.5 * ls).backward()
    # reference formulas
    X, S, FS = (t.detach().double().requires_grad_(True) for t in (xyz, sc, fs))
    vis = radii > 0
    relu = torch.nn.functional.relu
    if metric_xyz:
        rx = relu((X * FS[binding])[vis] - thr_x).norm(dim=1).mean() * lam_x
    else:
        rx = relu(X[vis].norm(dim=1) - thr_x).mean() * lam_x
    if lam_scale != 0:
        if metric_scale:
            rs = relu((torch.exp(S) * FS[binding])[vis] - thr_s).norm(dim=1).mean() * lam_scale
        else:
            rs = relu(torch.exp(S[vis]) - thr_s).norm(dim=1).mean() * lam_scale
    else:
        rs = torch.zeros((), dtype=torch.float64, device=dev)
    (3.0 * rx + 0.5 * rs).backward()
    assert int(cnt) == int(vis.sum())
    assert abs(float(lx) - float(rx)) <= 2e-6 * abs(float(rx)) + 1e-12
    assert abs(float(ls) - float(rs)) <= 2e-6 * abs(float(rs)) + 1e-12
    for got, ref, name in ((xyz.grad, X.grad, "xyz"), (sc.grad, S.grad, "scaling")):
        ref = torch.zeros_like(got, dtype=torch.float64) if ref is None else ref
        # d||relu(v - t)|| / dv = relu(v - t) / ||.|| is the DIRECTION of a vector that can be arbitrarily short (every
        # component within 1e-6 of the threshold): float32 exp vs the float64 reference then disagree on it, bounded
        # by the term's weight.  Elementwise gate + a handful of such splats.
        err = (got.double() - ref).abs()
        tol = 1e-4 * ref.abs() + 1e-6 * float(ref.abs().max())
        bad = int((err > tol).sum())
        assert bad <= max(6, int(2e-4 * err.numel())), f"{name}: {bad} entries beyond tolerance (worst {float(err.max()):.3e})"
        assert float(err.max()) <= 1.01 * float(ref.abs().max()) + 1e-12, name
        assert float(got[~vis].abs().sum()) == 0.0
    if metric_xyz or (metric_scale and lam_scale != 0):
        assert torch.allclose(fs.grad.double(), FS.grad, rtol=2e-4, atol=1e-10), "face_scaling"
    else:
        assert fs.grad is None


def test_binding_regularizers_on_a_rendered_frame_and_unbound():
    """With the radii of a real frame, and the plain-model form (no binding: non-metric terms only)."""
    import gaussianavatars_b200 as g

    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(3)
    xyz = torch.randn(1000, 3, generator=gen).to(dev).requires_grad_(True)
    sc = torch.randn(1000, 3, generator=gen).to(dev).requires_grad_(True)
    vis = (torch.rand(1000, generator=gen) > 0.3).to(dev)
    lx, ls = g.binding_regularizers(xyz, sc, vis)
    (lx + ls).backward()
    rx = torch.nn.functional.relu(xyz.detach()[vis].norm(dim=1) - 1.0).mean() * 1e-2
    assert abs(float(lx) - float(rx)) < 1e-7 and xyz.grad.shape == (1000, 3)
    # nothing visible: mean over an empty set is nan in the reference, and here
    lx0, _ = g.binding_regularizers(xyz.detach(), sc.detach(), torch.zeros(1000, dtype=torch.bool, device=dev))
    assert torch.isnan(lx0)
