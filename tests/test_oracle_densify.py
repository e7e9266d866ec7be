"""CPU: oracle/densify.py (the gather-plan restatement the CUDA op follows) against the REAL reference densification
run on CPU (tests/golden/make_golden_densify.py -> densify_vectors.npz)."""
import os

import numpy as np
import pytest

from oracle import densify as od

GOLD = os.path.join(os.path.dirname(__file__), "golden", "densify_vectors.npz")


def load_case(case):
    z = np.load(GOLD)
    params = {n: z[f"{case}_in_{n}"] for n in od.NAMES}
    state = {n: (z[f"{case}_in_{n}_exp_avg"], z[f"{case}_in_{n}_exp_avg_sq"]) for n in od.NAMES}
    stats = {k: z[f"{case}_in_{k}"] for k in ("xyz_gradient_accum", "denom", "max_radii2D")}
    bound = f"{case}_in_binding" in z.files
    extra = dict(binding=z[f"{case}_in_binding"].astype(np.int64), binding_counter=z[f"{case}_in_binding_counter"],
                 face_scaling=z[f"{case}_in_face_scaling"]) if bound else {}
    want = {n: z[f"{case}_out_{n}"] for n in od.NAMES}
    want_state = {n: (z[f"{case}_out_{n}_exp_avg"], z[f"{case}_out_{n}_exp_avg_sq"]) for n in od.NAMES}
    want_b = (z[f"{case}_out_binding"], z[f"{case}_out_binding_counter"]) if bound else (None, None)
    return params, state, stats, z[f"{case}_hyper"], z[f"{case}_noise"], extra, want, want_state, want_b, z


@pytest.mark.parametrize("case", ["bound", "plain"])
def test_gather_plan_reproduces_the_reference_densification(case):
    params, state, stats, hyper, noise, extra, want, want_state, want_b, z = load_case(case)
    out_p, out_s, b, c, P2 = od.densify_and_prune(params, state, stats, hyper, noise, **extra)
    assert P2 == want["xyz"].shape[0]
    pl = od.plan(params, stats, hyper, **extra)
    assert pl["clone"].sum() > 0 and pl["split"].sum() > 0 and (~pl["keep_orig"] & ~pl["split"]).sum() > 0, \
        "fixture must exercise clone, split and prune"
    for n in od.NAMES:
        if n in ("xyz", "scaling"):   # children: a bmm / exp-log chain evaluated in a different association
            # (3-term dot products with cancellation: the error scales with the terms, i.e. with the largest entry)
            assert np.allclose(out_p[n], want[n], rtol=2e-6, atol=1e-6 * float(np.abs(want[n]).max())), n
        else:
            assert np.array_equal(out_p[n], want[n]), n
        assert np.array_equal(out_s[n][0], want_state[n][0]) and np.array_equal(out_s[n][1], want_state[n][1]), n
    if want_b[0] is not None:
        assert np.array_equal(b, want_b[0]) and np.array_equal(c, want_b[1])
        assert (c > 0).all() or (z[f"{case}_in_binding_counter"] == 0).any()
    # the statistics come back zeroed at the new length (densification_postfix)
    for k in ("xyz_gradient_accum", "denom", "max_radii2D"):
        assert z[f"{case}_out_{k}"].shape[0] == P2 and not z[f"{case}_out_{k}"].any()
