"""GPU parity tests (run on the B200 box with -m gpu): the CUDA path through the public API / C-ABI against the
oracle (oracle/restated.py) on the same seeded inputs and against the committed reference-generated golden
vectors.  Tolerances: north_star's relative 1e-9 in FP64 (arm: 1e-6 against HEAD-generated goldens, whose own
dense inverse is the inaccurate party - SURVEY section 7); projection active sets bit-identical."""
import numpy as np
import pytest

from oracle import problems as P, restated as R

pytestmark = pytest.mark.gpu


def _gpu():
    import gpu_util
    return gpu_util


def _agree_fraction(a, b):
    return float(np.mean(np.all(a.reshape(a.shape[0], -1) == b.reshape(b.shape[0], -1), axis=1)))


def _check_admm(out, o, tol, utol):
    g = _gpu()
    assert np.array_equal(out["n_log"], o["n_log"]), "outer iteration counts differ"
    assert g.rel_logs(out["cost_log"], o["cost_log"]) < tol
    assert np.array_equal(out["admm_iters"], o["admm_iters"]), "ADMM iteration counts differ"
    assert np.array_equal(out["status"], o["status"])
    assert np.array_equal(out["admm_exit"], o["admm_exit"])
    assert _agree_fraction(out["alpha_idx"], o["alpha_idx"]) == 1.0, "line-search index sequences differ"
    assert np.abs(out["u"] - o["u"]).max() < utol
    assert np.abs(out["x"] - o["x"]).max() < utol
    assert np.abs(out["z_u"] - o["z_u"]).max() < utol
    assert np.abs(out["lam_u"] - o["lam_u"]).max() < utol
    m = ~np.isnan(o["res_log"])
    assert np.array_equal(np.isnan(out["res_log"]), ~m)
    assert np.max(np.abs(out["res_log"][m] - o["res_log"][m])) < 1e3 * utol


def _masks_identical(out, o, p, key, near=1e-12):
    """Active set of the last projection must be bit-identical except where the pre-projection value lies within
    rounding of a bound (reported separately)."""
    mo, mg = o["mask_" + key], out["mask_" + key]
    diff = mo != mg
    if not diff.any():
        return 0
    # elements allowed to differ: z within `near` of a bound on both sides
    lo, hi = p["lo_" + key], p["hi_" + key]
    z = o["z_" + key]
    nearb = (np.abs(z - lo) < near * np.maximum(1, np.abs(lo))) | (np.abs(z - hi) < near * np.maximum(1, np.abs(hi)))
    assert np.all(nearb[diff]), "clip masks differ away from the bounds"
    return int(diff.sum())


def test_car_ilqr_admm_vs_oracle():
    p = P.car_batch(96)
    out = _gpu().run_ilqr_admm(p)
    o = R.ilqr_admm(p)
    _check_admm(out, o, 1e-9, 1e-9)
    assert _masks_identical(out, o, p, "u") == 0
    assert np.all(np.abs(out["z_u"]) <= 0.5)          # z is feasible bit-exactly


def test_car_ilqr_admm_vs_reference_golden(golden):
    g = golden("car_ilqr_admm")
    p = P.car_batch(6)
    out = _gpu().run_ilqr_admm(p)
    ref = g["cost_log"]
    assert np.array_equal(out["n_log"], (~np.isnan(ref)).sum(1))
    cl = out["cost_log"][:, :ref.shape[1]]
    assert _gpu().rel_logs(cl, ref) < 1e-9
    assert np.abs(out["u"] - g["u"]).max() < 1e-9
    assert np.abs(out["x"] - g["x"]).max() < 1e-9


def test_car_stress_vs_oracle_and_golden(golden):
    """dt=0.03, theta0 in [0, 2pi): exercises the mod-2pi wrap and diverging line searches."""
    p = P.car_batch(32, stress=True)
    out = _gpu().run_ilqr_admm(p)
    o = R.ilqr_admm(p)
    assert np.array_equal(out["n_log"], o["n_log"])
    # chaotic problems amplify rounding differences: compare the common prefix of identical alpha sequences
    same = np.all(out["alpha_idx"].reshape(32, -1) == o["alpha_idx"].reshape(32, -1), axis=1)
    assert same.mean() >= 0.9
    assert _gpu().rel_logs(out["cost_log"][same], o["cost_log"][same]) < 1e-7
    g = golden("car_stress_ilqr_admm")
    ref = g["cost_log"]
    out3 = _gpu().run_ilqr_admm(P.car_batch(3, stress=True))
    assert _gpu().rel_logs(out3["cost_log"][:, :ref.shape[1]], ref) < 1e-7


def test_car_fixed_budget_vs_oracle():
    p = P.car_batch(64, I_o=6, I_a=5, L=20)
    out = _gpu().run_ilqr_admm(p, fixed_budget=True)
    o = R.ilqr_admm(p, fixed_budget=True)
    assert np.all(out["outer_iters"] == 6) and np.all(out["admm_iters"] == 5)
    assert _gpu().rel_logs(out["cost_log"], o["cost_log"]) < 1e-9
    assert np.abs(out["u"] - o["u"]).max() < 1e-9


def test_arm_ilqr_admm_vs_oracle():
    p = P.arm_batch(32)
    out = _gpu().run_ilqr_admm(p)
    o = R.ilqr_admm(p)
    # cost spans 3e6 -> 0.2 through cond~1e7 solves: 1e-9 relative to the running cost scale is not attainable
    # for the first iterates by any FP64 implementation; compare iterate-wise at 1e-7 and the converged cost at 1e-8
    assert np.array_equal(out["n_log"], o["n_log"])
    assert np.array_equal(out["admm_iters"], o["admm_iters"])
    print("arm iLQR-ADMM: max rel cost_log diff %.3e, final cost rel %.3e, max|du| %.3e, max|dx| %.3e" % (
        _gpu().rel_logs(out["cost_log"], o["cost_log"]), np.max(np.abs(out["cost"] - o["cost"]) / o["cost"]),
        np.abs(out["u"] - o["u"]).max(), np.abs(out["x"] - o["x"]).max()))
    # (measured: GPU vs oracle 4e-7 on the worst iterate; the unmodified reference itself is 3e-7..4e-7 away from the
    # oracle on these iterates, tests/test_oracle_golden.py::test_arm_ilqr_admm_matches_reference)
    assert _gpu().rel_logs(out["cost_log"], o["cost_log"]) < 2e-6
    assert np.max(np.abs(out["cost"] - o["cost"]) / o["cost"]) < 1e-8
    assert np.abs(out["u"] - o["u"]).max() < 1e-6
    assert np.abs(out["x"] - o["x"]).max() < 1e-7
    nd_u = _masks_identical(out, o, p, "u", near=1e-7)
    nd_x = _masks_identical(out, o, p, "x", near=1e-7)
    print("arm near-bound mask differences: u", nd_u, "x", nd_x)


def test_arm_ilqr_admm_vs_reference_golden(golden):
    g = golden("arm_ilqr_admm")
    out = _gpu().run_ilqr_admm(P.arm_batch(3))
    ref = g["cost_log"]
    assert np.array_equal(out["n_log"], (~np.isnan(ref)).sum(1))
    assert _gpu().rel_logs(out["cost_log"][:, :ref.shape[1]], ref) < 1e-6
    assert np.abs(out["u"] - g["u"]).max() < 5e-6


def test_car_ilqr_dp_vs_oracle_and_golden(golden):
    g = golden("car_ilqr_dp")
    p = P.car_batch(64)
    out = _gpu().run_ilqr_dp(p, int(g["max_iter"]), int(g["L"]))
    o = R.ilqr_dp(p, max_iter=int(g["max_iter"]), L=int(g["L"]))
    assert np.array_equal(out["n_log"], o["n_log"])
    assert np.array_equal(out["status"], o["status"])
    assert np.array_equal(out["alpha_idx"][:, :, 0], o["alpha_idx"]), "line-search index sequences differ"
    # unregularised iLQR amplifies rounding on a few hard problems (problem 28 stalls with "forward pass failed" at
    # cost 3.07; there the unmodified reference differs from the oracle by 2e-7, the GPU by 1e-8): require 1e-9 on
    # >= 95 % of the problems and 1e-6 on all
    m = ~np.isnan(o["cost_log"])
    rel = np.where(m, np.abs(out["cost_log"] - o["cost_log"]) / np.where(m, np.abs(o["cost_log"]), 1.0), 0.0).max(1)
    assert np.array_equal(np.isnan(out["cost_log"]), ~m)
    assert np.quantile(rel, 0.95) < 1e-9 and rel.max() < 1e-6
    du = np.abs(out["u"] - o["u"]).reshape(64, -1).max(1)
    dK = np.abs(out["K"] - o["K"]).reshape(64, -1).max(1) / np.abs(o["K"]).reshape(64, -1).max(1)
    print("plain iLQR car: quantiles of max|du| (50/90/100 %):", np.quantile(du, [0.5, 0.9, 1.0]),
          " of rel dK:", np.quantile(dK, [0.5, 0.9, 1.0]))
    assert np.quantile(du, 0.5) < 1e-9 and du.max() < 1e-5
    assert np.quantile(dK, 0.5) < 1e-9 and dK.max() < 1e-5
    ref = g["cost_log"]
    out4 = _gpu().run_ilqr_dp(P.car_batch(4), int(g["max_iter"]), int(g["L"]))
    assert _gpu().rel_logs(out4["cost_log"][:, :ref.shape[1]], ref) < 1e-9
    assert np.abs(out4["u"] - g["u"]).max() < 1e-9


def test_arm_ilqr_dp_vs_oracle():
    p = P.arm_batch(32)
    out = _gpu().run_ilqr_dp(p, 20, 25)
    o = R.ilqr_dp(p, max_iter=20, L=25)
    assert np.array_equal(out["n_log"], o["n_log"])
    m = ~np.isnan(o["cost_log"])
    rel0 = np.max(np.abs(out["cost_log"][m] - o["cost_log"][m]) / o["cost_log"][:, :1].repeat(21, 1)[m])
    print("arm plain iLQR: max cost_log diff relative to initial cost %.3e, final cost rel %.3e, max|du| %.3e" % (
        rel0, np.max(np.abs(out["cost"] - o["cost"]) / o["cost"]), np.abs(out["u"] - o["u"]).max()))
    assert rel0 < 1e-9
    assert np.max(np.abs(out["cost"] - o["cost"]) / o["cost"]) < 1e-8
    assert np.abs(out["u"] - o["u"]).max() < 1e-6


def test_di_lqt_admm_dp_vs_oracle_and_golden(golden):
    g = golden("di_lqt_admm_dp")
    p = P.di_batch(3)
    out = _gpu().run_lqt_admm_dp(p)
    assert np.array_equal(out["admm_iters"][:, 0], g["iters"])
    assert np.abs(out["x"] - g["x"]).max() < 1e-9
    assert np.abs(out["u"] - g["u"]).max() < 1e-9
    assert np.abs(out["K"] - g["K"]).max() / np.abs(g["K"]).max() < 1e-9
    p = P.di_batch(64)
    out = _gpu().run_lqt_admm_dp(p)
    o = R.lqt_admm_dp(p)
    assert np.array_equal(out["admm_iters"][:, 0], o["iters"])
    assert np.array_equal(out["admm_exit"][:, 0], o["exit_code"])
    assert np.abs(out["x"] - o["x"]).max() < 1e-9
    assert np.abs(out["u"] - o["u"]).max() < 1e-9
    assert np.abs(out["z_x"] - o["z_x"]).max() < 1e-9
    assert np.array_equal(out["mask_u"], o["mask_u"])
    assert np.isclose(np.abs(out["z_u"]).max(), 3.0) and np.abs(out["z_x"][:, :, 2:]).max() <= 0.6


def test_notebook_pin_lqt_admm_dp(golden):
    """'LQR and SLS with control bounds' notebook (n=2, m=1, N=100): ADMM_LQT_DP converges in 55 iterations at
    HEAD to the golden trajectory."""
    g = golden("notebook_pins")
    N = 100
    seq = np.zeros(N, dtype=np.int32); seq[-1] = 1
    p = dict(model="double_integrator", dt=0.01, N=N, n=2, m=1, zs=np.array([[0.0, 0.0], [1.0, 0.0]]),
             Qdiag=np.array([[0.0, 0.0], [1e6, 1e6]]), seq=seq, u_std=1e-2, x0=np.zeros((1, 2)),
             u0=np.zeros((N, 1)), lo_u=np.full((N, 1), -5.0), hi_u=np.full((N, 1), 5.0), lo_x=None, hi_x=None,
             rho_u=np.full((N, 1), 1e-1), rho_x=None, I_o=1, I_a=2000, L=1, tol=1e-4, alpha=1.0)
    out = _gpu().run_lqt_admm_dp(p)
    assert int(out["admm_iters"][0, 0]) == int(g["di_admm_dp_iters"])
    assert np.abs(out["u"][0] - g["di_admm_dp_u"]).max() < 1e-9
    assert np.abs(out["x"][0] - g["di_admm_dp_x"]).max() < 1e-9


def test_riccati_teacher_forced(golden):
    """isls_riccati_f64 on the reference's own inputs vs the reference's backward_pass_DP output."""
    import torch
    from isls_b200 import solver as S
    for name in ("car_backward_pass", "arm_backward_pass"):
        g = golden(name)
        t = lambda a: torch.as_tensor(a, device="cuda:0")[None].repeat(5, *([1] * a.ndim))
        K, k, bad = S.riccati(t(g["A"]), t(g["B"]), t(g["c"]), t(g["C"]))
        assert int(bad.sum()) == 0
        K, k = K.cpu().numpy(), k.cpu().numpy()
        for b in range(5):
            assert np.abs(K[b] - g["K"]).max() / np.abs(g["K"]).max() < 1e-10
            assert np.abs(k[b] - g["k"]).max() / np.abs(g["k"]).max() < 1e-10


@pytest.mark.parametrize("n,m", [(9, 3), (4, 2), (6, 3)])
def test_riccati_generic_ragged_batches_vs_oracle(n, m):
    """isls_riccati_f64 on random dense operators (incl. Cux) against the oracle's backward_pass: batches that do not
    fill the last 32-problem tile and horizons of both parities - the (9, 3) kernel fetches its A / B blocks with
    16-byte bulk copies from a source shifted by the parity of (problem * N + t)."""
    import torch
    from isls_b200 import solver as S
    rng = np.random.default_rng(100 * n + m)
    for B, N in ((37, 7), (64, 12), (5, 100), (33, 31)):
        A = np.eye(n) + 0.05 * rng.normal(size=(B, N, n, n))
        Bm = 0.1 * rng.normal(size=(B, N, n, m))
        c = rng.normal(size=(B, N, n + m))
        W = rng.normal(size=(B, N, n + m, n + m))
        C = W @ np.swapaxes(W, -1, -2) + np.eye(n + m)
        t = lambda a: torch.as_tensor(a, device="cuda:0")
        K, k, bad = S.riccati(t(A), t(Bm), t(c), t(C))
        assert int(bad.sum()) == 0
        Ko, ko, bo = R.backward_pass(A, Bm, c[..., :n], c[..., n:], C[..., :n, :n], C[..., n:, n:], Cux=C[..., n:, :n])
        assert not bo.any()
        eK = np.abs(K.cpu().numpy() - Ko).max() / np.abs(Ko).max()
        ek = np.abs(k.cpu().numpy() - ko).max() / np.abs(ko).max()
        assert eK < 1e-10 and ek < 1e-10, (n, m, B, N, eK, ek)


def test_linesearch_stage_teacher_forced():
    """isls_rollout_linesearch_f64: same x_nom, u_nom, du, reg -> per-candidate costs, argmin and winner rollout
    of the oracle (isls.py:468-477)."""
    import torch
    from isls_b200 import solver as S
    from oracle import models as M
    rng = np.random.default_rng(5)
    for p in (P.car_batch(40), P.arm_batch(40)):
        B, N, n, m = 40, p["N"], p["n"], p["m"]
        model = R._model_of(p)
        x_nom, u_nom = R.initial_rollout(p)
        du = rng.normal(0, 0.3, (B, N, m))
        reg_u = rng.normal(0, 0.3, (B, N, m))
        reg_x = rng.normal(0, 0.3, (B, N, n))
        al = R.alphas(p["L"])
        u_c = u_nom[:, None] + al[None, :, None, None] * du[:, None]
        x_c = R.rollout_open(model, x_nom[:, 0], u_c)
        zs = R._zs_b(p, B)
        costs = R.quad_cost(p, zs, x_c, u_c)
        if p["lo_x"] is not None:
            costs = costs + np.sum((x_c - reg_x[:, None]) ** 2 * p["rho_x"], axis=(-1, -2))
        costs = costs + np.sum((u_c - reg_u[:, None]) ** 2 * p["rho_u"], axis=(-1, -2))
        plan = S.Plan(p["model"], N, n, m, p["dt"], p["Qdiag"], p["seq"], p["u_std"], p["L"],
                      rho_x=p["rho_x"], lo_x=p["lo_x"], hi_x=p["hi_x"], rho_u=p["rho_u"], lo_u=p["lo_u"],
                      hi_u=p["hi_u"])
        sv = S.BatchSolver(plan, B, logs=False)
        c, best, xb, ub = sv.linesearch(x_nom, u_nom, du, zs, reg_x if p["lo_x"] is not None else None, reg_u)
        c, best, xb, ub = c.cpu().numpy(), best.cpu().numpy(), xb.cpu().numpy(), ub.cpu().numpy()
        assert np.max(np.abs(c - costs) / np.abs(costs)) < 1e-11
        ind = np.argmin(costs, axis=1)
        assert np.array_equal(best, ind)
        ar = np.arange(B)
        assert np.abs(xb - x_c[ar, ind]).max() < 1e-10
        assert np.abs(ub - u_c[ar, ind]).max() < 1e-13


def test_admm_project_dual_bit_exact():
    """isls_admm_project_dual_f64 vs numpy (admm.py:43-59 + np.clip): bit-identical z, lambda and clip masks,
    including relax != 1, ragged length, infinite bounds."""
    import torch
    from isls_b200 import solver as S
    rng = np.random.default_rng(7)
    for (B, ln, relax) in ((1, 1, 1.0), (3, 7, 1.0), (64, 200, 1.0), (17, 401, 1.6), (5, 900, 0.7)):
        x = rng.normal(0, 2, (B, ln)); z = rng.normal(0, 2, (B, ln)); lam = rng.normal(0, 1, (B, ln))
        lo = rng.normal(-1, 0.5, ln); hi = lo + rng.uniform(0, 2, ln)
        lo[::5] = -np.inf; hi[::7] = np.inf
        pre = relax * x + (1 - relax) * z + lam
        zn = np.clip(pre, lo, hi)
        r = x - zn
        ln_new = lam + r
        mask = (pre > hi).astype(np.int8) - (pre < lo).astype(np.int8)
        td = lambda a: torch.as_tensor(a, device="cuda:0").contiguous()
        zt, lt = td(z.copy()), td(lam.copy())
        prim, dual, mk = S.admm_project_dual(td(x), zt, lt, td(lo), td(hi), relax, want_mask=True)
        assert np.array_equal(zt.cpu().numpy(), zn)
        assert np.array_equal(lt.cpu().numpy(), ln_new)
        assert np.array_equal(mk.cpu().numpy(), mask)
        assert np.allclose(prim.cpu().numpy(), np.sum(r * r, axis=1), rtol=1e-13)
        assert np.allclose(dual.cpu().numpy(), np.sum((zn - z) ** 2, axis=1), rtol=1e-13)


def test_batch_composition_invariance_and_ragged():
    """Sharding property: a problem's result does not depend on which batch it is solved in (bit-for-bit), incl.
    ragged batch sizes (B not a multiple of the 32-problem tile) and B=1."""
    p = P.car_batch(77)
    full = _gpu().run_ilqr_admm(p)
    for idx in (np.arange(1), np.arange(40, 77), np.array([3, 50, 76])):
        sub = _gpu().run_ilqr_admm(P.subset(p, idx))
        for k in ("x", "u", "cost_log", "z_u", "status", "alpha_idx"):
            assert np.array_equal(sub[k], full[k][idx], equal_nan=True), k


def test_kernel_selection_regimes_agree_bitwise():
    """The launch rules pick other kernel forms by batch size (csrc/isls_kernels.cuh, launch_ff / launch_kpass): below
    513 tiles the deep-ring ff-pass with the Jacobian cache, from 513 to 1,184 tiles the shallow-ring form next to the
    small-batch K-pass, from 1,536 tiles the large-batch forms.  The same 96 problems embedded in a 96-, a 20,000- and a
    50,000-problem batch must come out bit for bit the same (the strong-scaling shards of one job take all three)."""
    I_o, I_a = 2, 3
    big = P.car_batch(50000, I_o=I_o, I_a=I_a, L=20)
    idx = np.random.default_rng(5).choice(20000, 96, replace=False)
    outs = []
    for B in (50000, 20000):
        o = _gpu().run_ilqr_admm(P.subset(big, np.arange(B)), fixed_budget=True, want_masks=False)
        outs.append({k: o[k][idx] for k in ("x", "u", "cost_log", "z_u")})
    o = _gpu().run_ilqr_admm(P.subset(big, idx), fixed_budget=True, want_masks=False)
    outs.append({k: o[k] for k in ("x", "u", "cost_log", "z_u")})
    for other in outs[1:]:
        for k in outs[0]:
            assert np.array_equal(outs[0][k], other[k], equal_nan=True), k


def test_full_size_properties():
    """BASELINE size (car, N=100, B=65,536), fixed budget: size-independent properties - determinism across
    runs, agreement of a random subsample with the oracle, feasibility of z, cost_log consistent with cost."""
    B = 65536
    p = P.car_batch(B, I_o=3, I_a=5, L=20)
    a = _gpu().run_ilqr_admm(p, fixed_budget=True, want_masks=False)
    b = _gpu().run_ilqr_admm(p, fixed_budget=True, want_masks=False)
    for k in ("x", "u", "cost_log"):
        assert np.array_equal(a[k], b[k], equal_nan=True)
    assert np.all(np.abs(a["z_u"]) <= 0.5)
    assert np.array_equal(a["cost"], a["cost_log"][:, 3])
    idx = np.random.default_rng(0).choice(B, 48, replace=False)
    o = R.ilqr_admm(P.subset(p, idx), fixed_budget=True)
    assert _gpu().rel_logs(a["cost_log"][idx], o["cost_log"]) < 1e-9
    assert np.abs(a["u"][idx] - o["u"]).max() < 1e-9


def test_api_surface_lqt_and_continuation(golden):
    """Remaining pieces of the reference surface on the hot path: SLS.solve / solve_dp / compute_cost, state
    persistence of iSLS between calls and iterate_once_dp (isls.py:336-374)."""
    import torch
    from isls_b200 import SLS, get_double_integrator_AB
    from oracle import models as M
    p = P.di_batch(4)
    p["x0"][:, :2] = np.array([[0.0, 0.0], [0.01, -0.02], [0.03, 0.0], [-0.02, 0.02]])
    s = SLS(4, 2, p["N"], batch=4)
    s.AB = get_double_integrator_AB(2, 2, p["dt"])
    s.set_quadratic_cost(p["zs"], p["Qdiag"], p["seq"], p["u_std"])
    x, u = s.solve(p["x0"], method="dp")
    K, k = s.solve_dp(x0=p["x0"])
    # oracle: unregularised Riccati + closed-loop rollout
    N, n, m = p["N"], 4, 2
    A, B = M.double_integrator_AB(2, 2, p["dt"])
    Qd = p["Qdiag"][p["seq"]]
    Ab, Bb = np.broadcast_to(A, (1, N, n, n)), np.broadcast_to(B, (1, N, n, m))
    zs_t = p["zs"][p["seq"]]
    Ko, ko, _ = R.backward_pass(Ab, Bb, (-2.0 * Qd * zs_t)[None], np.zeros((1, N, m)),
                                R._diag_embed(2.0 * Qd)[None], R._diag_embed(np.full((N, m), 2.0 * p["u_std"]))[None])
    assert np.abs(K.cpu().numpy() - Ko).max() / np.abs(Ko).max() < 1e-9
    xs, us = R.linear_rollout(np.broadcast_to(Ab, (4, N, n, n)), np.broadcast_to(Bb, (4, N, n, m)),
                              np.broadcast_to(Ko, (4, N, m, n)), np.broadcast_to(ko, (4, N, m)), dx0=p["x0"])
    assert np.abs(x.cpu().numpy().reshape(4, N, n) - xs).max() < 1e-9
    assert np.abs(u.cpu().numpy().reshape(4, N, m) - us).max() < 1e-8
    c = s.compute_cost(x.reshape(4, N, n), u.reshape(4, N, m)).cpu().numpy()
    co = R.quad_cost(p, R._zs_b(p, 4), xs[:, None], us[:, None])[:, 0]
    assert np.abs(c - co).max() / np.abs(co).max() < 1e-9
    # iSLS: two calls of max_iter=3 continue where the first stopped == one call of max_iter=6 (fixed budget)
    pc = P.car_batch(32)
    s1 = _gpu().make_isls(pc)
    s1.solve("car", max_iter=3, max_line_search_iter=20, fixed_budget=True)
    s1.solve("car", max_iter=3, max_line_search_iter=20, fixed_budget=True)
    s2 = _gpu().make_isls(pc)
    s2.solve("car", max_iter=6, max_line_search_iter=20, fixed_budget=True)
    assert torch.equal(s1.u_nom, s2.u_nom) and torch.equal(s1.cost, s2.cost)
    ok, K1, k1 = s2.iterate_once_dp(max_line_search=20)
    assert ok.dtype == torch.bool and K1.shape == (32, pc["N"], 2, 4)


def test_edge_cases_small_sizes_relaxation_and_state_bounds():
    """Edge cases the reference's semantics define: minimal horizon / single candidate / single problem, ADMM
    relaxation alpha != 1 (un-relaxed dual update, admm.py:48-52), state AND control bounds on the car (the
    trajectory-per-thread k_admm path), infinite bounds (no projection effect)."""
    g = _gpu()
    # (a) N = 3, L = 1, B = 1
    p = P.car_batch(1, N=3, I_o=3, I_a=2, L=1)
    out, o = g.run_ilqr_admm(p), R.ilqr_admm(p)
    assert np.array_equal(out["n_log"], o["n_log"]) and g.rel_logs(out["cost_log"], o["cost_log"]) < 1e-10
    assert np.abs(out["u"] - o["u"]).max() < 1e-10
    # (b) relaxation alpha = 1.6
    p = P.car_batch(40, I_o=8)
    p["alpha"] = 1.6
    out, o = g.run_ilqr_admm(p), R.ilqr_admm(p)
    assert np.array_equal(out["n_log"], o["n_log"]) and np.array_equal(out["alpha_idx"], o["alpha_idx"])
    assert g.rel_logs(out["cost_log"], o["cost_log"]) < 1e-9
    assert np.abs(out["lam_u"] - o["lam_u"]).max() < 1e-9 and np.array_equal(out["mask_u"], o["mask_u"])
    # (c) car with a speed limit (state bound on v) + control bounds: exercises k_admm and the x penalty terms
    p = P.car_batch(48, I_o=10)
    N = p["N"]
    lo_x, hi_x = np.full((N, 4), -np.inf), np.full((N, 4), np.inf)
    lo_x[:, 3], hi_x[:, 3] = -0.8, 0.8
    rho_x = np.zeros((N, 4)); rho_x[:, 3] = 5.0
    p.update(lo_x=lo_x, hi_x=hi_x, rho_x=rho_x)
    out, o = g.run_ilqr_admm(p), R.ilqr_admm(p)
    assert np.array_equal(out["n_log"], o["n_log"]) and np.array_equal(out["admm_iters"], o["admm_iters"])
    same = np.all(out["alpha_idx"].reshape(48, -1) == o["alpha_idx"].reshape(48, -1), axis=1)
    assert same.mean() > 0.95
    assert g.rel_logs(out["cost_log"][same], o["cost_log"][same]) < 1e-9
    assert np.abs(out["z_x"][same] - o["z_x"][same]).max() < 1e-9
    assert np.array_equal(out["mask_x"][same], o["mask_x"][same]) and np.array_equal(out["mask_u"][same], o["mask_u"][same])
    assert np.abs(out["z_x"][:, :, 3]).max() <= 0.8
    # (d) infinite bounds: projection is the identity, lambda stays 0, z = primal iterate
    p = P.car_batch(8, I_o=3)
    p["lo_u"][:], p["hi_u"][:] = -np.inf, np.inf
    out = g.run_ilqr_admm(p, fixed_budget=True)
    assert np.all(out["lam_u"] == 0.0) and np.array_equal(out["z_u"], out["u"]) and np.all(out["mask_u"] == 0)


def test_status_flags_nan_and_non_pd():
    """Per-problem failure reporting: a NaN initial state poisons that problem only (np.argmin picks the first NaN,
    isls.py:477 / NaN -> 1e5 in plain iLQR, isls.py:362); a non-PD Quu (the reference raises LinAlgError,
    isls.py:296) sets ISLS_ST_NON_PD and the problem stops with 'forward pass failed'."""
    g = _gpu()
    p = P.car_batch(34, I_o=3)
    p["x0"][5, 0] = np.nan
    out = g.run_ilqr_admm(p, fixed_budget=True)
    assert out["status"][5] & 32 and np.isnan(out["cost"][5])
    ok = np.ones(34, bool); ok[5] = False
    assert not np.isnan(out["cost"][ok]).any() and not (out["status"][ok] & 32).any()
    q = P.subset(p, np.arange(34)[ok])
    ref = g.run_ilqr_admm(q, fixed_budget=True)
    assert np.array_equal(out["u"][ok], ref["u"])                    # neighbours in the tile are untouched
    assert np.all(out["alpha_idx"][5] == 0)                          # first NaN wins the argmin
    # plain iLQR: NaN costs become 1e5 -> the step is rejected (cost is NaN, dcost < 0 is False)
    o2 = g.run_ilqr_dp(p, 3, 10)
    assert o2["status"][5] & 32 and o2["status"][5] & 2
    # non-PD: negative control weight
    p = P.car_batch(6, I_o=2)
    p["u_std"] = -5.0
    o3 = g.run_ilqr_dp(p, 2, 10)
    assert np.all(o3["status"] & 16) and np.all(o3["status"] & 2) and np.all(o3["n_log"] == 1)


def test_di_lqt_admm_batch_form_and_stage_api(golden):
    """SLS.ADMM_LQT_Batch (sls.py:250-294), SLS.solve_dp_ff (sls.py:168-202) and iSLS.rollout_DP (isls.py:310-334)
    of the reference's surface."""
    import torch
    from isls_b200 import Bound, SLS, get_double_integrator_AB
    g = golden("di_lqt_admm_batch")
    p = P.di_batch(3)
    s = SLS(p["n"], p["m"], p["N"], batch=3)
    s.AB = get_double_integrator_AB(p["m"], 2, p["dt"])
    s.set_quadratic_cost(p["zs"], p["Qdiag"], p["seq"], p["u_std"])
    x, u, log = s.ADMM_LQT_Batch(p["x0"], project_x=Bound(p["lo_x"], p["hi_x"]), project_u=Bound(p["lo_u"], p["hi_u"]),
                                 rho_x=p["rho_x"], rho_u=p["rho_u"], max_iter=p["I_a"], tol=p["tol"], log=True)
    its = s.last.admm_iters[:, 0].cpu().numpy()
    assert np.array_equal(its, g["iters"]), "ADMM_LQT_Batch iteration counts differ from the reference"
    assert np.abs(x.cpu().numpy().reshape(3, p["N"], -1) - g["x"]).max() < 1e-9
    assert np.abs(u.cpu().numpy().reshape(3, p["N"], -1) - g["u"]).max() < 1e-9
    o = R.lqt_admm_dp(p, batch_form=True)
    assert np.abs(u.cpu().numpy().reshape(3, p["N"], -1) - o["u"]).max() < 1e-9
    # solve_dp / solve_dp_ff (isls/sls.py:85-202) against the unmodified reference: unregularised and the regularised
    # (ADMM) form with diagonal Qr, Rr, incl. the return_Qs logs
    gd = golden("di_solve_dp_ff")
    s1 = SLS(p["n"], p["m"], p["N"])
    s1.AB = get_double_integrator_AB(p["m"], 2, p["dt"])
    s1.set_quadratic_cost(p["zs"], p["Qdiag"], p["seq"], p["u_std"])
    rel = lambda a, b: float(np.abs(a.cpu().numpy() - b).max() / max(np.abs(b).max(), 1e-300))
    K, k, Quu, Qui, Qux = s1.solve_dp(return_Qs=True)
    for a, nm in ((K, "K"), (k, "k"), (Quu, "Quu"), (Qui, "Quu_inv"), (Qux, "Qux")):
        assert rel(a, gd[nm]) < 1e-9, nm
    assert rel(s1.solve_dp_ff(K, Quu, Qux, Qui), gd["k_ff"]) < 1e-9
    Qr = np.stack([np.diag(q) for q in gd["qr"]])
    Rr = [np.diag(r) for r in gd["rr"]]
    Kr, kr, Quur, Quir, Quxr = s1.solve_dp(Qr=Qr, Rr=Rr, ur=gd["ur"], xr=gd["xr"], return_Qs=True)
    for a, nm in ((Kr, "Kr"), (kr, "kr"), (Quur, "Quur"), (Quir, "Quu_invr"), (Quxr, "Quxr")):
        assert rel(a, gd[nm]) < 1e-9, nm
    assert rel(s1.solve_dp_ff(Kr, Quur, Quxr, Quir, Qr=Qr, Rr=Rr, ur=gd["ur2"], xr=gd["xr2"]), gd["kr_ff"]) < 1e-9
    with pytest.raises(NotImplementedError):
        s1.solve_dp(Qr=Qr + 0.1, xr=gd["xr"])                          # dense Qr: SURVEY D10
    # rollout_DP: closed-loop rollouts of the line-search candidates around the nominal trajectory (arm: no angle
    # wrap, so the comparison is not at the mercy of a mod-2pi flip under feedback)
    pc = P.arm_batch(1)
    c = _gpu().make_isls(pc)
    c.batch, c.nb = None, 1
    c.solve("arm3", max_iter=3, max_line_search_iter=10, fixed_budget=True)
    al = torch.as_tensor(c.alphas[:4], device=c.k.device)
    xl, ul = c.rollout_DP(c.K, al[:, None, None] * c.k[None])
    model = R._model_of(pc)
    xs, us = R.rollout_closed(model, c.x_nom.cpu().numpy()[None], c.u_nom.cpu().numpy()[None],
                              c.K.cpu().numpy()[None], (al[:, None, None] * c.k[None]).cpu().numpy()[None])
    scale = max(1.0, np.abs(us).max())
    assert np.abs(xl.cpu().numpy() - xs[0]).max() < 1e-9 * scale and np.abs(ul.cpu().numpy() - us[0]).max() < 1e-9 * scale


def test_ragged_horizon_and_candidate_counts():
    """Horizons that are not a multiple of the line search's staged chunk (10 steps; ragged last bulk copy) and
    candidate counts that do not fill the CTA shapes (L = 33 -> 7 of 10 warps, L = 7 -> 2 warps)."""
    for N, L, B in ((37, 33, 5), (11, 7, 33), (101, 20, 3)):
        p = P.car_batch(B, N=N, I_o=4, I_a=3, L=L)
        out = _gpu().run_ilqr_admm(p, fixed_budget=True)
        o = R.ilqr_admm(p, fixed_budget=True)
        assert np.array_equal(out["alpha_idx"], o["alpha_idx"]), (N, L)
        assert _gpu().rel_logs(out["cost_log"], o["cost_log"]) < 1e-9, (N, L)
        assert np.abs(out["u"] - o["u"]).max() < 1e-9 and np.abs(out["z_u"] - o["z_u"]).max() < 1e-9, (N, L)


def test_host_admm_driver_vs_reference_golden(golden):
    """isls_b200.ADMM (the reference's generic driver, isls/admm.py:6-106, with the projection / dual update in the
    isls_admm_project_dual_f64 kernel) on the toy problem the reference's own ADMM() solved for the fixture: identical
    iteration counts and residual logs, z / lambda / primal iterates to 1e-12, for alpha = 1 and the over-relaxed 1.5."""
    import torch
    from isls_b200 import ADMM, Bound
    g = golden("admm_toy")
    a, b = torch.as_tensor(g["a"], device="cuda:0"), torch.as_tensor(g["b"], device="cuda:0")
    rho = float(g["rho"])
    for tag, alpha in (("a10", 1.0), ("a15", 1.5)):
        x, u, lx, lu, zx, zu, logs = ADMM(12, 8, lambda rx, ru: ((a + rho * rx) / (1 + rho), (b + rho * ru) / (1 + rho)),
                                          project_x=Bound(-0.5, 0.5), project_u=Bound(-1.0, 1.5), max_iter=200,
                                          alpha=alpha, tol=1e-6, return_lmb=True, log=True)
        ref = g[tag + "_logs"]
        assert len(logs) == len(ref), "iteration counts differ"
        assert np.allclose(np.array(logs), ref, rtol=1e-9, atol=1e-15)
        for t, nm in ((x, "x"), (u, "u"), (lx, "lx"), (lu, "lu"), (zx, "zx"), (zu, "zu")):
            assert np.abs(t.cpu().numpy() - g[tag + "_" + nm]).max() < 1e-12, nm


def test_lqt_general_lti_pair_vs_reference_golden(golden):
    """Base.AB takes any constant (A, B) (isls/base.py:98-119): LQT / LQT-ADMM (DP and batch form) with a coupled damped
    oscillator - not a double integrator - on the dense "lti" device model against the unmodified reference
    (tests/golden/lqt_lti.npz): identical ADMM iteration counts, trajectories and gains to 1e-9."""
    from isls_b200 import SLS, Bound
    g = golden("lqt_lti")
    N, n, m = 40, 4, 2
    s = SLS(n, m, N, batch=3)
    s.AB = [g["A"], g["B"]]
    assert s._dt is None                                                   # not the registered double integrator
    s.set_quadratic_cost(g["zs"], g["Qdiag"], g["seq"], 1e-2)
    x, u = s.solve(g["x0"], method="dp")
    assert np.abs(x.cpu().numpy().reshape(3, N, n) - g["x_unc"]).max() < 1e-9
    assert np.abs(u.cpu().numpy().reshape(3, N, m) - g["u_unc"]).max() < 1e-8
    kw = dict(project_x=Bound(g["lo_x"], g["hi_x"]), project_u=Bound(-1.0, 1.0), rho_x=np.diag([0.0, 0.0, 1.0, 1.0]),
              rho_u=1e-1, max_iter=600, tol=1e-4)
    x, u, K, k = s.ADMM_LQT_DP(g["x0"], **kw)
    assert np.array_equal(s.last.admm_iters[:, 0].cpu().numpy(), g["iters"]), "ADMM_LQT_DP iteration counts differ"
    assert np.abs(x.cpu().numpy().reshape(3, N, n) - g["x"]).max() < 1e-9
    assert np.abs(u.cpu().numpy().reshape(3, N, m) - g["u"]).max() < 1e-8
    assert np.abs(K.cpu().numpy() - g["K"]).max() / np.abs(g["K"]).max() < 1e-9
    x, u = s.ADMM_LQT_Batch(g["x0"], **kw)
    assert np.array_equal(s.last.admm_iters[:, 0].cpu().numpy(), g["iters_batch"]), "ADMM_LQT_Batch iteration counts differ"
    assert np.abs(x.cpu().numpy().reshape(3, N, n) - g["x_batch"]).max() < 1e-9
    assert np.abs(u.cpu().numpy().reshape(3, N, m) - g["u_batch"]).max() < 1e-8


def test_non_default_device_matches_device_0():
    """A solver on cuda:1 while cuda:0 is the current device: the plan's constant block, the workspace and the per-device
    shared-memory opt-ins must all follow the solver's device.  Needs two GPUs."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    torch.cuda.set_device(0)
    p = P.car_batch(40, I_o=4, I_a=3)
    a = _gpu().run_ilqr_admm(p, device="cuda:0")
    b = _gpu().run_ilqr_admm(p, device="cuda:1")
    for k in ("x", "u", "cost_log", "z_u", "alpha_idx"):
        assert np.array_equal(a[k], b[k], equal_nan=True), k
    q = P.arm_batch(40, I_o=3, I_a=3)                    # > 48 KB of dynamic shared memory (TMA ring of the arm)
    a = _gpu().run_ilqr_admm(q, device="cuda:0")
    b = _gpu().run_ilqr_admm(q, device="cuda:1")
    assert np.array_equal(a["u"], b["u"]) and np.array_equal(a["cost_log"], b["cost_log"], equal_nan=True)
    d = _gpu().run_lqt_admm_dp(P.di_batch(5, max_iter=60), device="cuda:1")
    d0 = _gpu().run_lqt_admm_dp(P.di_batch(5, max_iter=60), device="cuda:0")
    assert np.array_equal(d["u"], d0["u"])
    from isls_b200 import _lib, solver as S
    plan = S.Plan("car", p["N"], 4, 2, p["dt"], p["Qdiag"], p["seq"], p["u_std"], p["L"], device="cuda:1")
    with pytest.raises(_lib.IslsError):
        S.BatchSolver(plan, 8, "cuda:0")
