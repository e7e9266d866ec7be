"""-m gpu: densify_and_prune on the device (csrc/densify.cu) against
  (a) the REAL reference densification run on CPU (tests/golden/densify_vectors.npz, made by make_golden_densify.py), fed
      the same split noise;
  (b) the gather-plan oracle (oracle/densify.py, itself pinned to (a)) at 150k splats;
  (c) the model-level wrapper: parameters replaced inside the optimizer the way the reference does it."""
import numpy as np
import pytest
import torch

from oracle import densify as od
from tests.test_oracle_densify import load_case

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _to_dev(params, state, stats, extra):
    t = lambda a, dt=torch.float32: torch.from_numpy(np.ascontiguousarray(a)).to(DEV, dt)  # noqa: E731
    p = {n: t(params[n]) for n in od.NAMES}
    s = {n: (t(state[n][0]), t(state[n][1])) for n in od.NAMES}
    b = dict(binding=t(extra["binding"], torch.int32), binding_counter=t(extra["binding_counter"], torch.int32),
             face_scaling=t(extra["face_scaling"])) if extra else {}
    return p, s, t(stats["xyz_gradient_accum"]), t(stats["denom"]), b


def _check(out_p, out_s, b_out, c_out, want, want_state, want_b):
    for n in od.NAMES:
        got = out_p[n].cpu().numpy()
        if n in ("xyz", "scaling"):
            assert np.allclose(got, want[n], rtol=2e-6, atol=1e-6 * float(np.abs(want[n]).max())), n
            kept = (got == want[n]).all(axis=tuple(range(1, got.ndim)))
            assert kept.sum() > 0
        else:
            assert np.array_equal(got, want[n]), n
        assert np.array_equal(out_s[n][0].cpu().numpy(), want_state[n][0]), n + " exp_avg"
        assert np.array_equal(out_s[n][1].cpu().numpy(), want_state[n][1]), n + " exp_avg_sq"
    if want_b[0] is not None:
        assert np.array_equal(b_out.cpu().numpy(), want_b[0]) and np.array_equal(c_out.cpu().numpy(), want_b[1])


@pytest.mark.parametrize("case", ["bound", "plain"])
def test_matches_the_real_reference_run(case):
    import gaussianavatars_b200 as g

    params, state, stats, hyper, noise, extra, want, want_state, want_b, _ = load_case(case)
    p, s, accum, denom, b = _to_dev(params, state, stats, extra)
    screen = None if hyper[3] <= 0 else float(hyper[3])
    out_p, out_s, b_out, c_out, info = g.densify_arrays(p, s, accum, denom, hyper[0], hyper[1], hyper[2], screen, hyper[4],
                                                        noise=torch.from_numpy(noise).to(DEV), **b)
    assert info["P_out"] == want["xyz"].shape[0] and info["split_parents"] * 2 == noise.shape[0]
    _check(out_p, out_s, b_out, c_out, want, want_state, want_b)


def test_150k_splats_against_the_gather_plan_oracle():
    """BASELINE config-3 size.  Also the cases the small fixtures cannot hold: a face whose every splat is a prune
    candidate (must keep them all), empty Adam state, and a second call on the result."""
    import gaussianavatars_b200 as g
    from gaussianavatars_b200 import synthetic as syn

    P, F = 150_000, 10_144
    sp = syn.avatar_splats(P, n_faces=F, seed=3, sh_degree=3)
    gen = torch.Generator().manual_seed(5)
    params = {"xyz": sp["_xyz"], "rotation": sp["_rotation"], "scaling": sp["_scaling"], "opacity": sp["_opacity"],
              "f_dc": sp["_features_dc"], "f_rest": sp["_features_rest"]}
    params = {k: v.numpy() for k, v in params.items()}
    binding = sp["binding"].numpy().astype(np.int64)
    params["opacity"][binding == 7] = -9.0          # face 7: every splat below min_opacity -> all must survive
    counter = np.bincount(binding, minlength=F).astype(np.int32)
    fs = (torch.rand(F, 1, generator=gen) * 0.004 + 0.001).numpy()
    state = {n: (torch.randn(params[n].shape, generator=gen).numpy() * 0.01,
                 torch.rand(params[n].shape, generator=gen).numpy() * 1e-4) for n in od.NAMES}
    stats = {"xyz_gradient_accum": (torch.rand(P, 1, generator=gen) * 6e-4).numpy(),
             "denom": torch.randint(0, 3, (P, 1), generator=gen).float().numpy(), "max_radii2D": np.zeros(P, np.float32)}
    hyper = np.array([0.0002, 0.005, 0.1, 20.0, 0.01])
    pl = od.plan(params, stats, hyper, binding, counter, fs)
    S = int(pl["split"].sum())
    assert S > 1000 and pl["clone"].sum() > 1000
    noise = torch.randn(2 * S, 3, generator=gen)
    want_p, want_s, want_b, want_c, P2 = od.densify_and_prune(params, state, stats, hyper, noise.numpy(), binding, counter, fs)
    assert (want_b == 7).sum() >= (binding == 7).sum()
    extra = dict(binding=binding, binding_counter=counter, face_scaling=fs)
    p, s, accum, denom, b = _to_dev(params, state, stats, extra)
    out_p, out_s, b_out, c_out, info = g.densify_arrays(p, s, accum, denom, *hyper[:3], 20.0, hyper[4], noise=noise.to(DEV), **b)
    assert info["P_out"] == P2
    _check(out_p, out_s, b_out, c_out, want_p, want_s, (want_b, want_c))
    # no optimizer state at all (first densification of a fresh optimizer) + noise drawn by the op itself
    none = {n: (None, None) for n in od.NAMES}
    o2 = g.densify_arrays(p, none, accum, denom, *hyper[:3], 20.0, hyper[4], generator=torch.Generator(DEV).manual_seed(1), **b)
    assert o2[4]["P_out"] == P2 and all(v == (None, None) for v in o2[1].values())
    assert torch.equal(o2[0]["f_rest"], out_p["f_rest"]) and torch.equal(o2[2], b_out)


def test_model_wrapper_replaces_parameters_inside_the_optimizer():
    import gaussianavatars_b200 as g
    from types import SimpleNamespace

    params, state, stats, hyper, noise, extra, want, want_state, want_b, _ = load_case("bound")
    t = lambda a, dt=torch.float32: torch.from_numpy(np.ascontiguousarray(a)).to(DEV, dt)  # noqa: E731
    m = SimpleNamespace()
    groups = []
    for n in od.NAMES:
        prm = torch.nn.Parameter(t(params[n]))
        setattr(m, g.densify.ATTR[n], prm)
        groups.append({"params": [prm], "lr": 1e-3, "name": n})
    groups.append({"params": [torch.nn.Parameter(torch.zeros(5, device=DEV))], "lr": 1e-3, "name": "expr"})  # FLAME group: untouched
    m.optimizer = g.Adam(groups, lr=0.0, eps=1e-15)
    for n in od.NAMES:
        prm = getattr(m, g.densify.ATTR[n])
        m.optimizer.state[prm] = {"step": torch.tensor(3.0), "exp_avg": t(state[n][0]), "exp_avg_sq": t(state[n][1])}
    m.xyz_gradient_accum, m.denom, m.max_radii2D = t(stats["xyz_gradient_accum"]), t(stats["denom"]), t(stats["max_radii2D"])
    m.percent_dense = float(hyper[4])
    m.binding, m.binding_counter = t(extra["binding"], torch.int64), t(extra["binding_counter"], torch.int32)
    m.face_scaling = t(extra["face_scaling"])
    info = g.densify_and_prune(m, hyper[0], hyper[1], hyper[2], float(hyper[3]), noise=t(noise))
    P2 = want["xyz"].shape[0]
    assert info["P_out"] == P2 and m._xyz.shape[0] == P2 and m._features_rest.shape == want["f_rest"].shape
    assert m.binding.dtype == torch.int64 and np.array_equal(m.binding.cpu().numpy(), want_b[0])
    assert not m.xyz_gradient_accum.any() and m.xyz_gradient_accum.shape == (P2, 1) and m.max_radii2D.shape == (P2,)
    by_name = {gr["name"]: gr for gr in m.optimizer.param_groups}
    for n in od.NAMES:
        prm = by_name[n]["params"][0]
        assert prm is getattr(m, g.densify.ATTR[n]) and prm.requires_grad and prm.is_leaf
        st = m.optimizer.state[prm]
        assert float(st["step"]) == 3.0 and np.array_equal(st["exp_avg"].cpu().numpy(), want_state[n][0])
    assert len(m.optimizer.state) == 6 and by_name["expr"]["params"][0].shape == (5,)
    for n in od.NAMES:   # and the optimizer steps on the new arrays
        getattr(m, g.densify.ATTR[n]).grad = torch.ones_like(getattr(m, g.densify.ATTR[n]))
    m.optimizer.step()
    assert float(m.optimizer.state[m._xyz]["step"]) == 4.0
