"""GPU parity for the obstacle-avoidance state projection (composition inside the solver): iLQR-ADMM with
project_x = project_set_convex over two rotated infinity-norm shells - the "parking between two cars" notebook
(Car/Iterative LQR with state constraints.ipynb cells 18-20, README animation_state_bounds.gif)."""
import numpy as np
import pytest

from oracle import problems as P, restated as R

pytestmark = pytest.mark.gpu


def _gpu():
    import gpu_util
    return gpu_util


def _rel(a, b):
    m = ~np.isnan(b)
    assert np.array_equal(np.isnan(a), ~m), "cost_log lengths differ"
    return np.where(m, np.abs(a - b) / np.where(m, np.abs(b), 1.0), 0.0).max(1)


def test_parking_notebook_problem(golden):
    g = golden("parking_obstacles")
    p = P.parking_batch(1)
    out = _gpu().run_ilqr_admm(p)
    ref = g["nb_cost_log"]
    assert out["n_log"][0] == len(ref)
    assert abs(out["cost_log"][0, 1] - 2564.0110889491493) < 1e-9 * 2564.0      # notebook printout (cell 20)
    rel = np.abs(out["cost_log"][0, :len(ref)] - ref) / np.abs(ref)
    print("parking notebook problem: rel cost_log", rel.max(), " max|dx|", np.abs(out["x"][0] - g["nb_x"]).max())
    assert rel.max() < 1e-7
    assert np.abs(out["x"][0] - g["nb_x"]).max() < 1e-6 and np.abs(out["u"][0] - g["nb_u"]).max() < 1e-6


def test_parking_batch_vs_oracle_and_golden(golden):
    g = golden("parking_obstacles")
    B = 40
    p = P.parking_batch(B, N=200, dt=0.075)
    out = _gpu().run_ilqr_admm(p)
    o = R.ilqr_admm(p)
    assert np.array_equal(out["n_log"], o["n_log"]), "outer iteration counts differ"
    same_admm = np.all(out["admm_iters"] == o["admm_iters"], axis=1)
    same_inner = np.all(out["inner_iters"] == o["inner_iters"], axis=(1, 2))
    rel = _rel(out["cost_log"], o["cost_log"])
    dx = np.abs(out["x"] - o["x"]).reshape(B, -1).max(1)
    print("parking batch: identical ADMM counts %.2f, identical inner-projection counts %.2f; rel cost_log quantiles "
          "(50/90/100 %%)" % (same_admm.mean(), same_inner.mean()), np.quantile(rel, [0.5, 0.9, 1.0]), " max|dx|",
          np.quantile(dx, [0.5, 0.9, 1.0]))
    assert out["inner_iters"].max() == 15
    same_alpha = np.all(out["alpha_idx"] == o["alpha_idx"], axis=(1, 2))
    ok = same_admm & same_inner & same_alpha
    print("  identical ADMM / inner / alpha sequences: %.2f of the problems" % ok.mean())
    assert ok.mean() >= 0.6
    # This problem class amplifies rounding: the regularised LQ solve is ill-conditioned (the unmodified reference and
    # the numpy oracle differ by 9e-8 on the 5-problem fixture, tests/test_oracle_golden.py), and once the state
    # constraint is inactive the primal residual is round-off (1e-15) so ADMM's relative-change stop test
    # (admm.py:78-80) compares noise - a flipped stop sends the two runs to different iterates (both valid).
    # Contract tested here: (1) every problem agrees to 1e-9 on the first outer iteration (before any
    # amplification; measured 1e-15), (2) problems with identical decision sequences agree to 1e-8 in the median and 1e-4 at worst,
    # (3) where the sequences differ, the first differing ADMM stop is noise-decided (primal residual < 1e-9).
    pre = np.abs(out["cost_log"][:, :2] - o["cost_log"][:, :2]) / np.abs(o["cost_log"][:, :2])
    assert pre.max() < 1e-9
    assert np.median(rel[ok]) < 1e-8 and rel[ok].max() < 1e-4 and dx[ok].max() < 1e-4
    for b in np.nonzero(~same_admm)[0]:
        j = int(np.argmax(out["admm_iters"][b] != o["admm_iters"][b]))
        a = min(out["admm_iters"][b, j], o["admm_iters"][b, j]) - 1
        assert min(out["res_log"][b, j, a, 0], o["res_log"][b, j, a, 0]) < 1e-9, (b, j, a)
    ref = g["cost_log"]
    out5 = _gpu().run_ilqr_admm(P.parking_batch(5, N=200, dt=0.075))
    assert np.array_equal(out5["n_log"], (~np.isnan(ref)).sum(1))
    r5 = _rel(out5["cost_log"][:, :ref.shape[1]], ref)
    print("parking vs reference fixture: rel cost_log", r5, " max|dx|", np.abs(out5["x"] - g["x"]).max())
    # noise-decided ADMM stops (see above) send problem 0 of the fixture down another path after a few iterations:
    # all five agree on the first iterate, at least four of them to 1e-6 throughout
    first = np.abs(out5["cost_log"][:, 1] - ref[:, 1]) / np.abs(ref[:, 1])
    assert first.max() < 1e-9 and np.sum(r5 < 1e-6) >= 4
    # the projection returns the consensus variable of the inner ADMM (like the reference), which is feasible only up to
    # that ADMM's residual: the penetration depth must match the oracle's
    ob = p["obstacles"]

    def margin(zx):
        return np.min([(np.abs((zx[:, :, :2] - ob["centers"][k]) @ ob["W"][k].T).max(-1) - ob["lower"][k]).min(1)
                       for k in range(2)], axis=0)
    assert np.abs(margin(out["z_x"]) - margin(o["z_x"]))[ok].max() < 1e-4
    assert margin(out["z_x"]).min() > -0.2


def test_di_spherical_obstacles_lqt_path(golden):
    """LQT-ADMM (DP and batch form) with the spherical-obstacle state projection = project_set_convex + Dykstra over
    quadratic shells (Double integrator/LQR and SLS with spherical obstacle avoidance.ipynb cells 12-14).  ADMM does not
    converge on this non-convex set and amplifies rounding (oracle vs unmodified reference: 4e-14 after 10 iterations,
    7e-5 after 50, O(1) after 200 - tests/test_oracle_golden.py), so the iterates are compared on the first iterations;
    over the notebook's full budget the device must stay feasible and reach the same cost level."""
    g = golden("di_obstacles")
    p = P.di_obstacle_batch(4, max_iter=12, tol=1e-4)
    out = _gpu().run_lqt_admm_dp(p, fixed_budget=True)
    o = R.lqt_admm_dp(p, fixed_budget=True)
    assert np.abs(out["res_log"][:, 0] - g["dp_logs"][:, :12]).max() < 1e-9, "residual logs differ from the reference"
    assert np.abs(out["res_log"][:, 0] - o["res_log"]).max() < 1e-9
    assert np.abs(out["x"] - o["x"]).max() < 1e-9 and np.abs(out["z_x"] - o["z_x"]).max() < 1e-9
    its = out["inner_iters"][:, 0, :12]
    assert np.array_equal(its // 1000, o["inner_iters"][:, :, 0]) and np.array_equal(its % 1000, o["inner_iters"][:, :, 1])
    pb = P.di_obstacle_batch(4, max_iter=12, tol=1e-3)
    outb = _gpu().run_lqt_admm_dp(pb, fixed_budget=True, batch_form=True)
    assert np.abs(outb["res_log"][:, 0] - g["batch_logs"][:, :12]).max() < 1e-9
    # full notebook budget (500 iterations): same cost level as the reference's printed 2.701e-01, z outside both spheres
    pf = P.di_obstacle_batch(4, max_iter=500, tol=1e-4)
    outf = _gpu().run_lqt_admm_dp(pf)
    print("spherical obstacles, 500 iterations: cost gpu", outf["cost"], " reference", g["dp_cost"])
    assert np.all(np.abs(outf["cost"] - g["dp_cost"]) < 0.05 * g["dp_cost"])
    ob = pf["obstacles"]
    for k in range(2):
        r2 = 0.5 * np.sum((outf["z_x"][:, :, :2] - ob["centers"][k]) ** 2, axis=-1)
        assert np.all(r2 > ob["lower"][k] * (1 - 1e-3))
