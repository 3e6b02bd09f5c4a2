"""CPU, container only: the oracle (oracle/restated.py) side by side with the UNMODIFIED reference
(`isls` at HEAD under /root/reference, driven through oracle/ref_shim.py) on seeded problems that are NOT in
tests/golden - a live check that the committed fixtures are not the only inputs on which the two agree.
Skipped wherever the reference tree is absent (the GPU box)."""
import numpy as np
import pytest

from oracle import models as M, problems as P, ref_shim as S, restated as R

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not S.available(), reason="reference tree not present (container only)")]


def _Qs(p):
    return np.stack([np.diag(q) for q in p["Qdiag"]])


def _clip(lo, hi):
    lo, hi = lo.flatten(), hi.flatten()
    return lambda z: np.clip(z, lo, hi)


def test_car_ilqr_admm_fresh_seed():
    """isls.py:379-501 (dense batch LS at HEAD) vs the Riccati-form restatement, seed outside the fixtures."""
    p = P.car_batch(2, seed=987654, I_o=8)
    o = R.ilqr_admm(p)
    model = M.make_model(p["model"], dt=p["dt"])
    for b in range(2):
        s = S.make_isls(model, p["N"], p["zs"], _Qs(p), p["seq"], p["u_std"])
        S.init_nominal(s, p["x0"][b], p["u0"])
        r = S.run_ilqr_admm(s, model, project_u=_clip(p["lo_u"], p["hi_u"]), rho_u=float(p["rho_u"].flat[0]),
                            max_iter=p["I_o"], max_admm_iter=p["I_a"], max_line_search_iter=p["L"], tol=p["tol"])
        n = len(r["cost_log"])
        assert o["n_log"][b] == n
        assert np.max(np.abs(o["cost_log"][b, :n] - r["cost_log"]) / np.abs(r["cost_log"])) < 1e-9
        assert np.abs(o["u"][b] - r["u"]).max() < 1e-9
        assert np.abs(o["x"][b] - r["x"]).max() < 1e-9


def test_backward_pass_dp_teacher_forced():
    """isls.py:229-308 on random SPD stage costs with Cux != 0 along a random car trajectory."""
    rng = np.random.default_rng(424242)
    model = M.make_model("car", dt=0.1)
    n, m, N = 4, 2, 40
    p = P.car_batch(1, N=N)
    s = S.make_isls(model, N, p["zs"], _Qs(p), p["seq"], p["u_std"])
    S.init_nominal(s, rng.normal(size=n), 0.3 * rng.normal(size=(N, m)))
    A, B = model.get_AB(s.x_nom, s.u_nom)
    s.AB = A, B
    G = rng.normal(size=(N, n + m, n + m))
    C = G @ np.swapaxes(G, 1, 2) + 0.5 * np.eye(n + m)
    c = rng.normal(size=(N, n + m))
    with S.quiet():
        K, k = s.backward_pass_DP(C, c)
    Ko, ko = R.backward_pass(A[None], B[None], c[None, :, :n], c[None, :, n:], C[None, :, :n, :n],
                             C[None, :, n:, n:], C[None, :, n:, :n])[:2]
    assert np.max(np.abs(Ko[0] - K)) / np.max(np.abs(K)) < 1e-10
    assert np.max(np.abs(ko[0] - k)) / np.max(np.abs(k)) < 1e-10
