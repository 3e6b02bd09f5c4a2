"""GPU: batched row projections against the reference's own functions restated inline
(isls/projections.py) on random rows incl. the boundary cases (inside / outside / zero rows / t < 0 for the SOC)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ref_linear(x, a, l, u):                     # projections.py:30-43
    aTx = np.sum(x * a, axis=-1)
    aTa = np.sum(a * a, axis=-1) + 1e-30
    z = x.copy()
    c1, c2 = aTx > u, aTx < l
    z[c1] = z[c1] - (aTx - u)[c1, None] * (a / aTa)
    z[c2] = z[c2] - (aTx - l)[c2, None] * (a / aTa)
    return z


def _ref_quadratic(x, l, u):                     # projections.py:86-104
    z = x.copy()
    val = 0.5 * np.sum(x * x, axis=-1)
    c1, c2 = val > u, l > val
    z[c1] = x[c1] * np.sqrt(2 * u) / np.linalg.norm(x[c1], axis=-1)[:, None]
    z[c2] = x[c2] * np.sqrt(2 * l) / np.linalg.norm(x[c2], axis=-1)[:, None]
    return z


def _ref_square(x, l, u):                        # projections.py:252-262
    z = x.copy()
    j = np.argmax(np.abs(x), axis=-1)
    cond = np.where(np.linalg.norm(x, ord=np.inf, axis=-1) < l)
    z[(cond, j[cond])] = l * np.sign(x[(cond, j[cond])])
    return np.maximum(np.minimum(z, u), -u)


def test_row_projections_vs_numpy():
    import torch
    from isls_b200 import projections as Pj
    from oracle import restated as R
    rng = np.random.default_rng(11)
    for dim in (2, 3, 5):
        x = rng.normal(0, 1.5, (4000, dim))
        x[::17] *= 0.05
        td = torch.as_tensor(x, device="cuda:0")
        g = lambda t: t.cpu().numpy()
        lo, hi = rng.normal(-1, 0.3, dim), rng.normal(1, 0.3, dim)
        assert np.array_equal(g(Pj.project_bound_batch(td, lo, hi)), np.clip(x, lo, hi))
        a = rng.normal(0, 1, dim)
        assert np.abs(g(Pj.project_linear_batch(td, a, -0.4, 0.7)) - _ref_linear(x, a, -0.4, 0.7)).max() < 1e-14
        assert np.abs(g(Pj.project_quadratic_batch(td, 0.3, 2.0)) - _ref_quadratic(x, 0.3, 2.0)).max() < 1e-14
        c = rng.normal(0, 0.5, dim)
        assert np.abs(g(Pj.project_quadratic_batch(td, 0.3, 2.0, center=c)) - (_ref_quadratic(x - c, 0.3, 2.0) + c)).max() < 1e-14
        b = rng.normal(0, 0.5, dim)
        k = 0.5 * b @ b
        assert np.abs(g(Pj.project_quadratic_b_batch(td, b, 0.3, 2.0)) - (_ref_quadratic(x + b, 0.3 + k, 2.0 + k) - b)).max() < 1e-14
        assert np.abs(g(Pj.project_square_batch(td, 0.4, 1.2)) - _ref_square(x, 0.4, 1.2)).max() < 1e-15
        assert np.abs(g(Pj.project_square_batch(td, 0.4, 1.2, center=c)) - (_ref_square(x - c, 0.4, 1.2) + c)).max() < 1e-15
        zz, tt = R.project_soc_unit_batch(x[:, :-1], x[:, -1])
        assert np.abs(g(Pj.project_soc_unit_batch(td)) - np.concatenate([zz, tt[:, None]], 1)).max() < 1e-15
        nrm = np.linalg.norm(x, axis=-1, keepdims=True)
        assert np.abs(g(Pj.project_unit_ball_batch(td)) - np.where(nrm <= 1, x, x / nrm)).max() < 1e-15
    # SOC batch semantics: [3, 0, -1] -> [0, 0, 0]
    out = Pj.project_soc_unit_batch(torch.tensor([[3.0, 0.0, -1.0], [0.5, 0.0, 1.0], [3.0, 0.0, 1.0]], device="cuda:0",
                                                 dtype=torch.float64)).cpu().numpy()
    assert np.array_equal(out[0], [0, 0, 0]) and np.array_equal(out[1], [0.5, 0, 1.0]) and np.allclose(out[2], [2, 0, 2])


def test_parameterised_projections_vs_reference_golden(golden):
    """project_multilinear, project_affine, project_soc (inner ADMM, identical iterates up to rounding) and
    project_block_lower_triangular (isls/projections.py:46-68, 163-232, 277-286) against outputs of the unmodified
    reference on the same seeded rows (tests/golden/projections_ex.npz)."""
    import torch
    from isls_b200 import projections as Pj
    g = golden("projections_ex")
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device="cuda:0")
    ml = Pj.project_multilinear_batch(t(g["x"]), g["A"], g["l"], g["u"]).cpu().numpy()
    assert np.abs(ml - g["ml"]).max() < 1e-12
    Ax = ml @ g["A"].T                                         # rows that were outside now sit on the violated bound
    assert np.all(Ax < g["u"] + 1e-9) and np.all(Ax > g["l"] - 1e-9)
    aff = Pj.project_affine_batch(t(g["x"]), g["a"], 0.7, -0.4, 0.6).cpu().numpy()
    assert np.abs(aff - g["aff"]).max() < 1e-13
    soc, its = Pj.project_soc_batch(t(g["z0"]), g["As"], g["bs"], rho=2.0, max_iter=100, tol=1e-6, want_iters=True)
    assert np.abs(soc.cpu().numpy() - g["soc"]).max() < 1e-10 and 1 <= its <= 100
    soc1 = Pj.project_soc_batch(t(g["z0"][3:4]), g["As"], g["bs"], rho=2.0, max_iter=100, tol=1e-6).cpu().numpy()
    assert np.abs(soc1[0] - g["soc1"]).max() < 1e-10          # a single row stops on its own residual (differs from row 3 of the batch call)
    blt = Pj.project_block_lower_triangular(t(g["Z"]), 4, 2, 6).cpu().numpy()
    assert np.array_equal(blt, g["blt"])


def test_generic_project_set_convex_vs_reference_golden(golden):
    """project_set_convex (isls/projections.py:289-374) over three sets of different kinds (box, quadratic shell with a
    centre, second-order cone) against the unmodified reference on the same 50 rows."""
    import torch
    from isls_b200 import projections as Pj
    g = golden("set_convex")
    x0 = torch.as_tensor(g["x0"], device="cuda:0")
    out, its = Pj.project_set_convex_batch(
        x0, [g["A0"], g["A1"], g["A2"]], [g["b0"], g["b1"], g["b2"]],
        [("bound", g["lo"], g["hi"]), ("quadratic", 0.05, 0.8, g["c1"]), ("soc_unit",)], rho=2.0, max_iter=150,
        threshold=1e-6, want_iters=True)
    assert np.abs(out.cpu().numpy() - g["out"]).max() < 1e-9 and 1 <= its <= 150
