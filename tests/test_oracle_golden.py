"""CPU: the oracle (oracle/restated.py) against the committed reference-generated golden vectors
(tests/golden/*.npz, produced by tests/golden/make_golden.py from the unmodified reference at HEAD)."""
import numpy as np

from oracle import problems as P, restated as R


def _rel(a, b):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))


def _check_logs(o, g, tol):
    for b in range(g["x"].shape[0]):
        ref = g["cost_log"][b]
        ref = ref[~np.isnan(ref)]
        assert o["n_log"][b] == len(ref), "iteration count differs from the reference"
        assert _rel(o["cost_log"][b, :len(ref)], ref) < tol


def test_car_ilqr_admm_matches_reference(golden):
    g = golden("car_ilqr_admm")
    p = P.car_batch(6)
    assert np.array_equal(p["x0"], g["x0"])
    o = R.ilqr_admm(p)
    _check_logs(o, g, 1e-9)
    assert np.abs(o["u"] - g["u"]).max() < 1e-9          # tolerance: north_star rel 1e-9 (|u| <= 0.5..1)
    assert np.abs(o["x"] - g["x"]).max() < 1e-9
    # active set: clipped controls sit exactly on the bound in z
    assert np.all(np.abs(o["z_u"]) <= 0.5)


def test_car_stress_distribution_matches_reference(golden):
    """dt=0.03 / theta0 in [0,2pi): hits the mod-2pi wrap; both sides must misbehave identically."""
    g = golden("car_stress_ilqr_admm")
    o = R.ilqr_admm(P.car_batch(3, stress=True))
    _check_logs(o, g, 1e-8)


def test_arm_ilqr_admm_matches_reference(golden):
    """HEAD's explicit inverse of an ill-conditioned l_side limits agreement to ~1e-6 here."""
    g = golden("arm_ilqr_admm")
    o = R.ilqr_admm(P.arm_batch(3))
    _check_logs(o, g, 1e-6)
    assert np.abs(o["u"] - g["u"]).max() < 5e-6
    assert np.abs(o["x"] - g["x"]).max() < 1e-6


def test_car_ilqr_dp_matches_reference(golden):
    g = golden("car_ilqr_dp")
    o = R.ilqr_dp(P.car_batch(4), max_iter=int(g["max_iter"]), L=int(g["L"]))
    _check_logs(o, g, 1e-9)
    assert np.abs(o["u"] - g["u"]).max() < 1e-9


def test_arm_ilqr_dp_matches_reference(golden):
    g = golden("arm_ilqr_dp")
    o = R.ilqr_dp(P.arm_batch(2), max_iter=int(g["max_iter"]), L=int(g["L"]))
    # the first iterates drop the cost from 3e6 to ~4 through a cond~1e7 solve (dposv in the reference, LU here):
    # intermediate costs agree to 1e-11 of the initial cost, the converged ones to 1e-8 relative
    for b in range(2):
        ref = g["cost_log"][b]
        ref = ref[~np.isnan(ref)]
        assert o["n_log"][b] == len(ref)
        assert np.abs(o["cost_log"][b, :len(ref)] - ref).max() < 1e-11 * ref[0]
        assert abs(o["cost_log"][b, len(ref) - 1] - ref[-1]) < 1e-8 * ref[-1]
    assert np.abs(o["u"] - g["u"]).max() < 1e-6


def _bp(g):
    A, B, C, c = g["A"], g["B"], g["C"], g["c"]
    n = A.shape[-1]
    K, k, bad = R.backward_pass(A[None], B[None], c[None, :, :n], c[None, :, n:], C[None, :, :n, :n],
                                C[None, :, n:, n:], Cux=C[None, :, n:, :n])
    assert not bad.any()
    return K[0], k[0]


def test_backward_pass_teacher_forced(golden):
    for name in ("car_backward_pass", "arm_backward_pass"):
        g = golden(name)
        K, k = _bp(g)
        assert np.abs(K - g["K"]).max() / np.abs(g["K"]).max() < 1e-10
        assert np.abs(k - g["k"]).max() / np.abs(g["k"]).max() < 1e-10


def test_di_lqt_admm_dp_matches_reference(golden):
    g = golden("di_lqt_admm_dp")
    o = R.lqt_admm_dp(P.di_batch(3))
    assert np.array_equal(o["iters"], g["iters"])
    assert np.abs(o["x"] - g["x"]).max() < 1e-9
    assert np.abs(o["u"] - g["u"]).max() < 1e-9
    assert np.abs(o["K"] - g["K"]).max() / np.abs(g["K"]).max() < 1e-9
    assert np.abs(o["res_log"][np.arange(3), g["iters"] - 1] - g["last_res"]).max() < 1e-9
    # both constraint families are active at the optimum
    assert np.isclose(np.abs(o["z_u"]).max(), 3.0) and np.isclose(np.abs(o["z_x"][:, :, 2:]).max(), 0.6)


def test_notebook_pins(golden):
    """Known answers printed in the reference notebooks: the DP LQT-ADMM run of
    'LQR and SLS with control bounds' (n=2,m=1,N=100, |u|<=5, rho_u=0.1, tol=1e-4)."""
    g = golden("notebook_pins")
    assert float(g["di_lqt_max_u"]) == 6.06051888764695
    assert float(g["di_lqt_last_pos"]) == 0.9999876316133441
    assert int(g["di_admm_batch_iters"]) == 20          # "converged at iteration 19"
    assert float(g["di_admm_batch_max_u"]) == 5.000018035934772
    N = 100
    seq = np.zeros(N, dtype=np.int32); seq[-1] = 1
    p = dict(model="double_integrator", dt=0.01, N=N, n=2, m=1, zs=np.array([[0.0, 0.0], [1.0, 0.0]]),
             Qdiag=np.array([[0.0, 0.0], [1e6, 1e6]]), seq=seq, u_std=1e-2, x0=np.zeros((1, 2)),
             u0=np.zeros((N, 1)), lo_u=np.full((N, 1), -5.0), hi_u=np.full((N, 1), 5.0), lo_x=None, hi_x=None,
             rho_u=np.full((N, 1), 1e-1), rho_x=None, I_o=1, I_a=2000, L=1, tol=1e-4, alpha=1.0)
    o = R.lqt_admm_dp(p)
    assert int(o["iters"][0]) == int(g["di_admm_dp_iters"])
    assert np.abs(o["u"][0] - g["di_admm_dp_u"]).max() < 1e-9
    assert np.abs(o["x"][0] - g["di_admm_dp_x"]).max() < 1e-9


def _sls_case(g, tag, pos_dim, N, dt, u_std=1e-2, Qf=1e6):
    from oracle import models as M
    n, m = 2 * pos_dim, pos_dim
    A, B = M.double_integrator_AB(pos_dim, 2, dt)
    Qt = np.zeros((N, n)); Qt[-1] = Qf
    tg = g[tag + "_targets"]
    xd = np.zeros((len(tg), N, n)); xd[:, -1, :pos_dim] = tg
    return A, B, Qt, xd.reshape(len(tg), -1), [a for a in g[tag + "_A_"]], [b for b in g[tag + "_b_"]]


def test_sls_admm_matches_reference(golden):
    """SLS.solve_sls + ADMM_SLS (SOC chance constraints): iteration counts, residual logs, d_u and the robust
    Phi_u columns against the unmodified reference; includes the notebook printout 'can't improve anymore at
    iteration 25', residual 8.50e-13 3.77e-01."""
    g = golden("sls_admm")
    for tag, pos_dim, N, dt in (("nb", 1, 100, 0.01), ("c4", 2, 50, 1.0 / 50)):
        A, B, Qt, xd, A_, b_ = _sls_case(g, tag, pos_dim, N, dt)
        o = R.admm_sls(A, B, N, Qt, xd, 1e-2, A_, b_, 1e2, max_iter=50, alpha=1.0, tol=1e-3, inner_rho=1e1,
                       inner_max_iter=100, inner_threshold=1e-3)
        assert np.array_equal(o["iters"], g[tag + "_iters"])
        for b in range(len(xd)):
            it = o["iters"][b]
            assert np.allclose(o["logs"][b], g[tag + "_logs"][b, :it], rtol=1e-6, atol=1e-15)
        assert np.abs(o["du0"] - g[tag + "_du0"]).max() < 1e-9
        assert np.abs(o["du"] - g[tag + "_du"]).max() < 1e-9
        c = pos_dim + 1
        assert np.abs(o["phi_u"][:, :, :c - 1] - g[tag + "_phic"]).max() < 1e-9
        # shared feedback part: the reference's Woodbury down-date chain (base.py:32-50) on cond(L) ~ 1e10 limits
        # agreement of PHI_U to ~1e-6 relative
        assert np.abs(o["PHI_U"] - g[tag + "_PHI0"]).max() / np.abs(g[tag + "_PHI0"]).max() < 1e-5
    assert int(g["nb_iters"][0]) == 26 and "%.2e %.2e" % tuple(g["nb_last"][0]) == "8.50e-13 3.77e-01"


def test_mc_rollouts_match_reference(golden):
    """get_trajectory_batch / dp / sls (noise-free) of SLS (double integrator) and iSLS (car) vs the reference."""
    from oracle import models as M
    g = golden("mc_rollouts")
    di = M.make_model("double_integrator", nb_dim=2, dt=0.05)
    for mode, K, k in (("dp", g["di_K"], g["di_k"]), ("sls", g["di_Ks"], g["di_ks"]), ("batch", None, g["di_us"])):
        x, u = R.mc_rollout(di, mode, g["di_x0"], K, k, 30)
        assert np.abs(x - g["di_%s_x" % mode]).max() < 1e-12 and np.abs(u - g["di_%s_u" % mode]).max() < 1e-12
    car = M.make_model("car", dt=0.1)
    x, u = R.mc_rollout(car, "dp", g["car_x0"], g["car_K"], g["car_k"], 25)
    assert np.abs(x - g["car_dp_x"]).max() < 1e-12 and np.abs(u - g["car_dp_u"]).max() < 1e-12
    x, u = R.mc_rollout(car, "sls", g["car_x0"], g["car_Ks"], g["car_ks"], 25, g["car_x_nom"], g["car_u_nom"])
    assert np.abs(x - g["car_sls_x"]).max() < 1e-12 and np.abs(u - g["car_sls_u"]).max() < 1e-12


def test_tutorial_tassa_pseudo_huber_vs_reference_golden(golden):
    """Tutorial problem (Tassa car + pseudo-Huber cost) - oracle vs the unmodified reference's
    solve(method='dp') and ilqr_admm (analytic get_AB / get_Cs callbacks), identical iteration counts."""
    g = golden("tutorial_tassa")
    p = P.tassa_batch(3, N=int(g["N"]))
    o = R.ilqr_dp(p, max_iter=100, L=40)
    ref = g["cost_log_dp"]
    assert np.array_equal(o["n_log"], (~np.isnan(ref)).sum(1))
    m = ~np.isnan(ref)
    assert np.max(np.abs(o["cost_log"][:, :ref.shape[1]][m] - ref[m]) / np.abs(ref[m])) < 1e-9
    assert np.abs(o["u"] - g["u_dp"]).max() < 1e-9 and np.abs(o["x"] - g["x_dp"]).max() < 1e-9
    o = R.ilqr_admm(p)
    ref = g["cost_log_admm"]
    assert np.array_equal(o["n_log"], (~np.isnan(ref)).sum(1))
    m = ~np.isnan(ref)
    assert np.max(np.abs(o["cost_log"][:, :ref.shape[1]][m] - ref[m]) / np.abs(ref[m])) < 1e-9
    assert np.abs(o["u"] - g["u_admm"]).max() < 1e-9 and np.abs(o["x"] - g["x_admm"]).max() < 1e-9


def _tassa_fullsize(g):
    p = P.tassa_batch(1, N=int(g["N"]))
    p["x0"] = g["x0"].copy()
    assert np.array_equal(p["u0"], g["u0"])
    return p


def test_tutorial_fullsize_vs_reference_golden(golden):
    """The Tutorial / 'Replicate of control-limited ddp car example' problem at the notebooks' own size (N = 500,
    T = 15 s, x0 = (1, 1, 3pi/2, 0)): oracle vs the unmodified reference, dp (60 iterations) and ADMM (45)."""
    g = golden("tutorial_tassa_n500")
    p = _tassa_fullsize(g)
    for tag, o in (("dp", R.ilqr_dp(p, max_iter=100, L=40)), ("admm", R.ilqr_admm(p))):
        ref = g["cost_log_" + tag][0]
        assert o["n_log"][0] == len(ref), tag
        rel = np.max(np.abs(o["cost_log"][0, :len(ref)] - ref) / np.abs(ref))
        du = np.abs(o["u"] - g["u_" + tag]).max()
        print("tutorial N=500", tag, "rel cost_log", rel, "max|du|", du)
        assert rel < 1e-8 and du < 1e-6, tag


def test_tassa_jacobian_finite_differences():
    """The hand-derived Jacobian of the Tassa car (the notebook uses autograd) against central differences."""
    from oracle import models as M
    mdl = M.make_model("tassa_car", dt=0.03)
    rng = np.random.default_rng(0)
    x = rng.normal(0, 1, (16, 4))
    x[:, 3] *= 3
    u = rng.normal(0, 0.4, (16, 2))
    A, B = mdl.get_AB(x, u)
    eps = 1e-6
    for i in range(4):
        d = np.zeros(4)
        d[i] = eps
        assert np.allclose((mdl.f(x + d, u) - mdl.f(x - d, u)) / (2 * eps), A[:, :, i], atol=1e-8)
    for j in range(2):
        d = np.zeros(2)
        d[j] = eps
        assert np.allclose((mdl.f(x, u + d) - mdl.f(x, u - d)) / (2 * eps), B[:, :, j], atol=1e-8)


def test_parking_obstacle_sets_vs_reference_golden(golden):
    """iLQR-ADMM with the state projection onto obstacle sets (project_set_convex over two rotated
    infinity-norm shells, Car/Iterative LQR with state constraints.ipynb cells 18-20) - oracle vs the unmodified
    reference: the notebook's own problem (known first iterate 2564.0110889491493 from the cell-20 output), five start
    states whose paths cross the obstacles, and the projection closure alone on random trajectories."""
    g = golden("parking_obstacles")
    assert abs(g["nb_cost_log"][1] - 2564.0110889491493) < 1e-9 * 2564.0          # HEAD reproduces the printout
    p = P.parking_batch(1)
    o = R.ilqr_admm(p)
    ref = g["nb_cost_log"]
    assert o["n_log"][0] == len(ref)
    assert np.max(np.abs(o["cost_log"][0, :len(ref)] - ref) / np.abs(ref)) < 1e-8
    assert np.abs(o["x"][0] - g["nb_x"]).max() < 1e-8 and np.abs(o["u"][0] - g["nb_u"]).max() < 1e-8
    p = P.parking_batch(5, N=200, dt=0.075)
    assert np.array_equal(p["x0"], g["x0"])
    o = R.ilqr_admm(p)
    # the state constraint is inactive for long stretches: the primal residual is then round-off (1e-15) and ADMM's
    # relative-change stop test (admm.py:78-80) is decided by noise, so single ADMM stops differ between arithmetics
    # (same situation as ADMM_SLS); the iterates agree to 1e-7, the outer iteration counts exactly
    _check_logs(o, g, 1e-6)
    assert np.abs(o["x"] - g["x"]).max() < 1e-5 and np.abs(o["u"] - g["u"]).max() < 1e-5
    assert o["inner_iters"].max() == 15                    # the inner projection ADMM is really exercised
    proj = R.obstacle_projections(p["obstacles"])
    ob = p["obstacles"]
    for q, ref in zip(g["proj_in"], g["proj_out"]):
        z, _ = R.project_set_convex_rows(q, proj, ob["rho"], ob["max_iter"], ob["threshold"])
        assert np.abs(z - ref).max() < 1e-12


def test_isls_admm_riccati_form_vs_reference_golden(golden):
    """robust nonlinear iSLS-ADMM (isls.py:503-712) - the Riccati-form restatement (one K-pass +
    (dim+1) feed-forward passes / linear rollouts instead of the dense (N m)^2 inverse) against the unmodified
    reference on the 3-DoF arm with SOC chance constraints: identical iteration counts, cost log 1e-9, d_u / Phi_u 1e-9."""
    g = golden("arm_isls_admm")
    B = g["x0"].shape[0]
    p = P.arm_robust_batch(B)
    assert np.array_equal(p["x0"], g["x0"])
    assert abs(g["cost_log"][0, 0] - 6775.068343357641) < 1e-9 * 6775.0      # notebook pin (initial cost)
    o = R.isls_admm(p)
    _check_logs(o, g, 1e-9)
    assert np.abs(o["u"] - g["u"]).max() < 1e-9 and np.abs(o["x"] - g["x"]).max() < 1e-9
    N, m = p["N"], p["m"]
    assert np.abs(o["d_u"].reshape(B, N * m) - g["du"]).max() < 1e-9
    assert np.abs(o["phi_u"].reshape(B, N * m, -1) - g["phi_u"]).max() < 1e-9 * max(1.0, np.abs(g["phi_u"]).max())
    assert np.abs(o["u"]).max() < 6.0                       # the nominal respects the (tightened) bound


def test_isls_admm_state_projection_vs_reference_golden(golden):
    """isls_admm with project_x (isls.py:556-559, 571-572, 631-638, 648-650): robust bounds on two joint velocities of
    the arm, with and without the control-side projection, against the unmodified reference (closure in
    oracle/ref_shim.run_isls_admm).  The state-side-only case is the cheap-control regime (R = 1e-4, no Rr) where the
    four-term value recursion fails and the closed-loop form is needed (backward_pass(joseph=True))."""
    g = golden("arm_isls_admm_x")
    p = P.arm_robust_x_batch(1)
    o = R.isls_admm(p)
    ref = g["cost_log"][0]
    n_ref = int((~np.isnan(ref)).sum())
    assert o["n_log"][0] == n_ref
    assert np.abs(o["cost_log"][0, :n_ref] - ref[:n_ref]).max() < 1e-9 * np.abs(ref[:n_ref]).max()
    N, m = p["N"], p["m"]
    assert np.abs(o["x"][0] - g["x"][0]).max() < 1e-9 and np.abs(o["u"][0] - g["u"][0]).max() < 1e-9
    assert np.abs(o["phi_u"][0].reshape(N * m, -1) - g["phi_u"][0]).max() < 1e-9 * max(1.0, np.abs(g["phi_u"][0]).max())
    p1 = P.arm_robust_x_batch(1, project_u=False)
    o1 = R.isls_admm(p1)
    r1 = g["xonly_cost_log"]
    assert o1["n_log"][0] == len(r1)
    assert np.abs(o1["cost_log"][0, :len(r1)] - r1).max() < 1e-8 * np.abs(r1).max()
    assert np.abs(o1["x"][0] - g["xonly_x"]).max() < 1e-7 and np.abs(o1["u"][0] - g["xonly_u"]).max() < 1e-5
    # the four-term recursion on the first linearisation of this problem: 10 % off in the first controls
    model = R._model_of(p1)
    xn, un = R.initial_rollout(p1)
    A, Bm = model.get_AB(xn, un)
    Qd = np.asarray(p1["Qdiag"], float)[p1["seq"]]
    Cxx = R._diag_embed(2.0 * (Qd + p1["robust"]["x"]["rho_x"]))[None]
    Cuu = R._diag_embed(np.full((N, m), 2.0 * p1["u_std"]))[None]
    z = np.zeros((1, N, p1["n"])), np.zeros((1, N, m))
    K4 = R.backward_pass(A, Bm, z[0], z[1], Cxx, Cuu)[0]
    Kj = R.backward_pass(A, Bm, z[0], z[1], Cxx, Cuu, joseph=True)[0]
    assert np.abs(K4[0, 0] - Kj[0, 0]).max() > 1e-3 * np.abs(Kj[0, 0]).max()


def test_lqt_admm_batch_form_vs_reference_golden(golden):
    """SLS.ADMM_LQT_Batch (isls/sls.py:250-294; SURVEY 8d C1: "converged at iteration 580"): the Riccati-form
    restatement with batch-form last control and the unconstrained warm start against the unmodified reference."""
    g = golden("di_lqt_admm_batch")
    p = P.di_batch(3)
    o = R.lqt_admm_dp(p, batch_form=True)
    assert np.array_equal(o["iters"], g["iters"])
    assert R.lqt_admm_dp(P.di_batch(1), batch_form=True)["iters"][0] == 581     # HEAD: "converged at iteration 580"
    assert np.abs(o["x"] - g["x"]).max() < 1e-9 and np.abs(o["u"] - g["u"]).max() < 1e-9
    last = o["res_log"][np.arange(3), o["iters"] - 1]
    assert np.abs(last - g["last_res"]).max() < 1e-9


def test_di_spherical_obstacles_vs_reference_golden(golden):
    """LQT-ADMM with the spherical-obstacle state projection (project_set_convex + Dykstra over quadratic
    shells; Double integrator/LQR and SLS with spherical obstacle avoidance.ipynb cells 12-14).  The constraint set is
    non-convex and ADMM never converges on it (the notebook prints "Max iteration reached"): rounding differences grow
    from 1e-14 to O(1) within ~100 iterations even between two numpy implementations, so the iterates are compared on
    the first iterations and the projection closure separately."""
    g = golden("di_obstacles")
    assert abs(g["dp_cost"][0] - 2.701e-01) < 5e-5            # notebook printout (cell 14)
    p = P.di_obstacle_batch(4, max_iter=12, tol=1e-4)
    assert np.array_equal(p["x0"], g["x0"])
    o = R.lqt_admm_dp(p, fixed_budget=True)
    assert np.abs(o["res_log"] - g["dp_logs"][:, :12]).max() < 1e-9
    o = R.lqt_admm_dp(P.di_obstacle_batch(4, max_iter=12, tol=1e-3), fixed_budget=True, batch_form=True)
    assert np.abs(o["res_log"] - g["batch_logs"][:, :12]).max() < 1e-9
    ob = p["obstacles"]
    for q, ref in zip(g["proj_in"], g["proj_out"]):
        z, i1, i2 = R.project_positions_spheres(q, ob)
        assert np.abs(z - ref).max() < 1e-12 and i1 == 5 and 1 <= i2 <= 51


def test_admm_sls_state_projection_vs_reference_golden(golden):
    """SLS.ADMM_SLS with project_u and project_x (sls.py:319-454; Double integrator/LQR and SLS with state bounds.ipynb
    cells 16-17): oracle vs the unmodified reference, incl. the notebook's printed residuals 1.37e-04 / 1.80e-01."""
    from oracle import models as M
    g = golden("sls_state_bounds")
    assert abs(g["logs"][-1, 0] - 1.37e-04) < 5e-7 and abs(g["logs"][-1, 1] - 1.80e-01) < 5e-4 and len(g["logs"]) == 100
    p = P.sls_state_bounds_problem()
    mdl = M.make_model("double_integrator", nb_dim=1, dt=p["dt"])
    xd = p["zs"][p["seq"]].reshape(-1)
    o = R.admm_sls(mdl.A, mdl.B, p["N"], np.zeros((p["N"], p["n"])), xd, p["u_std"], p["As"], p["bs_u"], p["rho_u"],
                   max_iter=p["max_iter"], alpha=1.0, tol=p["tol"], inner_rho=p["inner_rho"],
                   inner_max_iter=p["inner_max_iter"], inner_threshold=p["inner_threshold"], x_rows=p["x_rows"],
                   rho_x=p["rho_x"])
    assert o["iters"][0] == 100
    assert np.abs(o["logs"][0] - g["logs"]).max() < 1e-8
    assert np.abs(o["du"][0] - g["du"]).max() < 1e-9 and np.abs(o["phi_u"][0] - g["phi_u"]).max() < 1e-8
