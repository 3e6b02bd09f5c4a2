"""CPU: host-side logic of the mirror that needs no device - descriptors, rho expansion, plugin-slot type checks,
sharding arithmetic.  (Everything that computes runs in libisls_b200.so and is covered by the -m gpu tests.)"""
import numpy as np
import pytest

from isls_b200 import Bound, ObstacleSets, PseudoHuberCost, SetConvexSOC, iSLS, SLS, configs


def test_pseudo_huber_descriptor_matches_tutorial_config():
    p = configs.tassa_batch(1, N=30)
    a = PseudoHuberCost(cu=p["Rdiag"], cx=[1e-3, 1e-3], px=[0.1, 0.1], cf=[0.1, 0.1, 1.0, 0.3],
                        pf=[0.01, 0.01, 0.01, 1.0]).arrays(4, 30)
    assert np.array_equal(a["Ws"], p["Qdiag"]) and np.array_equal(a["Ps"], p["Hp"])
    assert np.array_equal(a["Ws_b"], p["Qdiag_b"]) and np.array_equal(a["Ps_b"], p["Hp_b"])
    assert np.array_equal(a["seq"], p["seq"]) and np.array_equal(a["zs"], p["zs"])


def test_obstacle_sets_descriptor():
    ob = configs.parking_batch(1)["obstacles"]
    d = ObstacleSets(ob["centers"], ob["W"], ob["lower"], ob["upper"], ob["rho"], ob["max_iter"], ob["threshold"])
    assert np.allclose(d.W_inv, ob["W_inv"]) and d.as_dict()["kind"] == "square"
    assert d.key() == ObstacleSets(ob["centers"], ob["W"], ob["lower"], ob["upper"], ob["rho"], ob["max_iter"],
                                   ob["threshold"]).key()
    # the rectangle interiors are the forbidden sets: their centres map to the origin of the W frame
    for k in range(2):
        assert np.allclose((ob["centers"][k] - ob["centers"][k]) @ ob["W"][k].T, 0.0)


def test_soc_descriptor_shapes():
    rb = configs.arm_robust_batch(1)["robust"]
    s = SetConvexSOC(rb["As"], rb["bs"], rho=rb["inner_rho"], max_iter=rb["inner_max_iter"],
                     threshold=rb["inner_threshold"])
    assert s.As.shape == (2, 5, 4) and s.bs.shape == (2, 5)
    with pytest.raises(AssertionError):
        SetConvexSOC([np.zeros((5, 4))], [np.zeros(4)])


def test_plugin_slots_reject_python_callables():
    s = iSLS(4, 2, 20, batch=2)
    with pytest.raises(TypeError):
        s.forward_model = lambda x, u: x
    with pytest.raises(TypeError):
        s.cost_function = lambda x, u: 0.0
    s.forward_model = ("car", {"dt": 0.1})
    with pytest.raises(TypeError):
        s._check_get_AB("arm3")
    with pytest.raises(TypeError):
        s._check_get_Cs(lambda x, u: None)
    s._check_get_Cs("analytic")
    s._check_get_AB("car")


def test_rho_expansion_matches_reference_semantics():
    """compute_Rr_Qr (isls/base.py:55-79): float, [dim, dim] and [N, dim, dim] inputs -> per-step diagonals."""
    s = iSLS(4, 2, 10)
    Qr, Rr = s.compute_Rr_Qr(np.diag([0.0, 0.0, 1.0, 1.0]), 1e-2)
    assert Qr.shape == (10, 4) and np.array_equal(Qr[3], [0.0, 0.0, 1.0, 1.0])
    assert Rr.shape == (10, 2) and np.all(Rr == 1e-2)
    r3 = np.zeros((10, 4, 4))
    r3[:, :2, :2] = np.eye(2) * 0.1
    Qr, _ = s.compute_Rr_Qr(r3, None)
    assert np.array_equal(Qr[0], [0.1, 0.1, 0.0, 0.0])
    with pytest.raises(Exception):
        s.compute_Rr_Qr(np.ones((4, 4)), None)            # dense rho: the reference's own expression needs it diagonal


def test_bound_expand_and_legacy_names():
    b = Bound([-0.5, -2.0], [0.5, 2.0]).expand(7, 2)
    assert b[0].shape == (7, 2) and np.all(b[1][:, 1] == 2.0)
    assert iSLS.set_cost_variables is iSLS.set_quadratic_cost
    assert hasattr(iSLS, "solve_ilqr") and hasattr(iSLS, "isls_admm") and hasattr(iSLS, "rollout_DP")
    assert hasattr(SLS, "ADMM_LQT_Batch") and hasattr(SLS, "solve_dp_ff") and hasattr(SLS, "ADMM_SLS")
