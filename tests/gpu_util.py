"""Helpers shared by the GPU parity tests: run an oracle problem dict (oracle/problems.py) through the public
isls_b200 API."""
import numpy as np

from isls_b200 import Bound, SLS, iSLS


def make_isls(p, device="cuda:0"):
    B = p["x0"].shape[0]
    s = iSLS(p["n"], p["m"], p["N"], batch=B, device=device)
    kw = {"dt": p["dt"]}
    s.forward_model = (p["model"], kw)
    if p.get("cost", "quadratic") == "pseudo_huber":
        s.set_pseudo_huber_cost(p["zs"], p["Qdiag"], p["Hp"], p["seq"], p["Rdiag"], p.get("Qdiag_b"), p.get("Hp_b"))
    else:
        s.set_quadratic_cost(p["zs"], p["Qdiag"], p["seq"], p["u_std"])
    s.set_initial(p["x0"], p["u0"])
    return s


def run_ilqr_admm(p, fixed_budget=False, want_masks=True, device="cuda:0"):
    s = make_isls(p, device)
    kw = {}
    if p.get("obstacles") is not None:
        from isls_b200 import ObstacleSets
        ob = p["obstacles"]
        kw.update(project_x=ObstacleSets(ob["centers"], ob["W"], ob["lower"], ob["upper"], ob["rho"], ob["max_iter"],
                                         ob["threshold"]), rho_x=p["rho_x"])
        want_masks = False
    if p.get("lo_x") is not None:
        kw.update(project_x=Bound(p["lo_x"], p["hi_x"]), rho_x=p["rho_x"])
    if p.get("lo_u") is not None:
        kw.update(project_u=Bound(p["lo_u"], p["hi_u"]), rho_u=p["rho_u"])
    out = s.ilqr_admm(max_iter=p["I_o"], max_admm_iter=p["I_a"], max_line_search_iter=p["L"], tol=p["tol"],
                      alpha=p.get("alpha", 1.0), fixed_budget=fixed_budget, want_masks=want_masks, **kw)
    return {k: v.cpu().numpy() for k, v in out.items()}


def run_ilqr_dp(p, max_iter, L, tol_fun=1e-5, fixed_budget=False, device="cuda:0"):
    s = make_isls(p, device)
    out = s.solve(p["model"], max_iter=max_iter, max_line_search_iter=L, tol_fun=tol_fun, fixed_budget=fixed_budget)
    return {k: v.cpu().numpy() for k, v in out.items()}


def run_lqt_admm_dp(p, fixed_budget=False, device="cuda:0", batch_form=False):
    from isls_b200 import get_double_integrator_AB
    B = p["x0"].shape[0]
    s = SLS(p["n"], p["m"], p["N"], batch=B, device=device)
    s.AB = get_double_integrator_AB(p["m"], 2, p["dt"])
    s.set_quadratic_cost(p["zs"], p["Qdiag"], p["seq"], p["u_std"])
    kw = {}
    if p.get("obstacles") is not None:
        from isls_b200 import ObstacleSets
        ob = p["obstacles"]
        kw.update(project_x=ObstacleSets(ob["centers"], None, ob["lower"], ob["upper"], ob["rho"], ob["max_iter"],
                                         ob["threshold"], kind=ob["kind"], dykstra_max_iter=ob["dykstra_max_iter"],
                                         dykstra_tol=ob["dykstra_tol"]), rho_x=p["rho_x"])
    if p.get("lo_x") is not None:
        kw.update(project_x=Bound(p["lo_x"], p["hi_x"]), rho_x=p["rho_x"])
    if p.get("lo_u") is not None:
        kw.update(project_u=Bound(p["lo_u"], p["hi_u"]), rho_u=p["rho_u"])
    fn = s.ADMM_LQT_Batch if batch_form else s.ADMM_LQT_DP
    fn(p["x0"], max_iter=p["I_a"], tol=p["tol"], alpha=p.get("alpha", 1.0), fixed_budget=fixed_budget,
       want_masks=p.get("obstacles") is None, **kw)
    return {k: v.cpu().numpy() for k, v in s.last.items()}


def rel_logs(a, b):
    m = ~np.isnan(b)
    assert np.array_equal(np.isnan(a), np.isnan(b)), "cost_log lengths differ"
    return np.max(np.abs(a[m] - b[m]) / np.maximum(np.abs(b[m]), 1e-300))


def run_isls_admm(p, fixed_budget=False, device="cuda:0"):
    from isls_b200 import SetConvexSOC
    rb = p["robust"]
    s = make_isls(p, device)
    soc = False if rb.get("As") is None else SetConvexSOC(rb["As"], rb["bs"], rho=rb["inner_rho"],
                                                          max_iter=rb["inner_max_iter"], threshold=rb["inner_threshold"])
    kw = {}
    if rb.get("u_unprojected"):
        soc = False
    if rb.get("x"):
        from isls_b200 import SetConvexSOCComponents
        rx = rb["x"]
        kw.update(project_x=SetConvexSOCComponents(rx["comps"], rb["As"], rx["bs"], rho=rb["inner_rho"],
                                                   max_iter=rb["inner_max_iter"], threshold=rb["inner_threshold"]),
                  rho_x=rx["rho_x"])
    s.isls_admm(rb["dim"], p["model"], project_u=soc, max_admm_iter=p["I_a"], k_max=p["I_o"], max_line_search=p["L"],
                rho_u=rb["rho_u"], alpha=p.get("alpha", 1.0), threshold=p["tol"], fixed_budget=fixed_budget, **kw)
    return {k: v.cpu().numpy() for k, v in s.last.items()}
