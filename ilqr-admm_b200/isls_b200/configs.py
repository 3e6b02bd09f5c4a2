"""Seeded synthetic problem batches C1..C5 of SURVEY.md 8(d) (BASELINE.json `configs`).

The same arrays are fed to the CUDA path (bench.py, tests), to the oracle and to the reference shim (golden
generation), so this module contains no solver logic - only problem data (numpy).

A problem batch is a plain dict:
  model      "car" | "arm3" | "double_integrator",  dt
  N, n, m
  zs[k,n], Qdiag[k,n], seq[N] int32, u_std     quadratic via-point cost (base.py:81-89)
  x0[B,n], u0[N,m]                              initial state per problem, shared initial control guess
  lo_u/hi_u[N,m], lo_x/hi_x[N,n] or None        box bounds (+-inf = unconstrained element)
  rho_u[N,m] or None, rho_x[N,n] or None        diagonal ADMM penalty weights (compute_Rr_Qr, base.py:55-79)
  I_o, I_a, L, tol, alpha                       budgets / tolerances (isls.py:379-381 keywords)
"""
import numpy as np

INF = np.inf


def _seed(idx):
    return 1234 + idx


def car_batch(B, N=100, dt=0.1, seed=_seed(2), I_o=20, I_a=5, L=20, tol=1e-3, stress=False):
    """C2 / C5: simple kinematic car, control bounds |u|<=0.5, rho_u = 10
    (weights: notebooks/Car/Iterative LQR with control constraints.ipynb cells 8, 18, 20)."""
    rng = np.random.default_rng(seed)
    n, m = 4, 2
    x0 = np.zeros((B, n))
    x0[:, 0] = rng.uniform(-2, 2, B)
    x0[:, 1] = rng.uniform(-2, 2, B)
    if stress:     # notebook-like distribution: dt=0.03, target theta 0, theta0 in [0, 2pi): hits the mod wrap
        x0[:, 2] = rng.uniform(0, 2 * np.pi, B)
        target = np.zeros(n)
    else:
        x0[:, 2] = rng.uniform(np.pi / 2, 3 * np.pi / 2, B)
        target = np.array([0.0, 0.0, np.pi, 0.0])
    zs = np.stack([np.zeros(n), target])
    Qdiag = np.stack([np.zeros(n), np.full(n, 1e2)])
    seq = np.zeros(N, dtype=np.int32)
    seq[-1] = 1
    return dict(name="car", model="car", dt=(0.03 if stress else dt), N=N, n=n, m=m, zs=zs, Qdiag=Qdiag, seq=seq,
                u_std=1e-2, x0=x0, u0=np.zeros((N, m)),
                lo_u=np.full((N, m), -0.5), hi_u=np.full((N, m), 0.5), lo_x=None, hi_x=None,
                rho_u=np.full((N, m), 1e1), rho_x=None, I_o=I_o, I_a=I_a, L=L, tol=tol, alpha=1.0)


def arm_batch(B, N=100, dt=0.01, seed=_seed(3), I_o=20, I_a=10, L=5, tol=1e-4):
    """C3: planar 3-DoF arm with state and control bounds
    (notebooks/3DoF robot/State and control bound constraints.ipynb cells 12, 15, 22-24)."""
    rng = np.random.default_rng(seed)
    n, m = 9, 3
    q0 = np.array([np.pi / 3, -np.pi / 2, -np.pi / 4]) + rng.normal(0.0, 0.1, (B, 3))
    a1 = q0[:, 0]
    a2 = a1 + q0[:, 1]
    a3 = a2 + q0[:, 2]
    fk = np.stack([np.cos(a1) + np.cos(a2) + np.cos(a3), np.sin(a1) + np.sin(a2) + np.sin(a3), np.zeros(B)], -1)
    x0 = np.concatenate([q0, np.zeros_like(q0), fk], axis=-1)          # state [q, qdot, p_ee], planar 3R unit links
    target = np.array([0, 0, 0, 0, 0, 0, 1.5, 1.0, 0.0])
    zs = np.stack([np.zeros(n), target])
    Qdiag = np.stack([np.zeros(n), np.array([0, 0, 0, 1e6, 1e6, 1e6, 0, 1e6, 0.0])])
    seq = np.zeros(N, dtype=np.int32)
    seq[-1] = 1
    lo_x = np.full((N, n), -INF)
    hi_x = np.full((N, n), INF)
    lo_x[:, 3:6], hi_x[:, 3:6] = -1.5, 1.5
    lo_x[-1, 6], hi_x[-1, 6] = 0.5, 1.0
    rho_x = np.zeros((N, n))
    rho_x[:, 3:6] = 1e-2
    rho_x[-1, 6] = 1e1
    return dict(name="arm3", model="arm3", dt=dt, N=N, n=n, m=m, zs=zs, Qdiag=Qdiag, seq=seq, u_std=1e-4,
                x0=x0, u0=np.ones((N, m)),
                lo_u=np.full((N, m), -6.0), hi_u=np.full((N, m), 6.0), lo_x=lo_x, hi_x=hi_x,
                rho_u=np.full((N, m), 1e-3), rho_x=rho_x, I_o=I_o, I_a=I_a, L=L, tol=tol, alpha=1.0)


def di_batch(B=1, N=50, seed=_seed(1), max_iter=2000, tol=1e-4, batched_targets=False):
    """C1: LQT-ADMM double integrator (n=4, m=2), state + control bounds
    (structure of notebooks/Double integrator/LQR and SLS with state bounds.ipynb)."""
    rng = np.random.default_rng(seed)
    n, m = 4, 2
    dt = 1.0 / N
    target = np.array([0.5, 0.4, 0.0, 0.0])
    zs = np.stack([np.zeros(n), target])
    Qdiag = np.stack([np.zeros(n), np.full(n, 1e3)])
    seq = np.zeros(N, dtype=np.int32)
    seq[-1] = 1
    x0 = np.zeros((B, n))
    if B > 1:
        x0[:, :2] = rng.normal(0.0, 0.02, (B, 2))
    lo_x = np.full((N, n), -INF)
    hi_x = np.full((N, n), INF)
    lo_x[:, 2:], hi_x[:, 2:] = -0.6, 0.6
    rho_x = np.zeros((N, n))
    rho_x[:, 2:] = 1.0
    return dict(name="double_integrator", model="double_integrator", dt=dt, N=N, n=n, m=m, zs=zs, Qdiag=Qdiag,
                seq=seq, u_std=1e-4, x0=x0, u0=np.zeros((N, m)),
                lo_u=np.full((N, m), -3.0), hi_u=np.full((N, m), 3.0), lo_x=lo_x, hi_x=hi_x,
                rho_u=np.full((N, m), 1e-2), rho_x=rho_x, I_o=1, I_a=max_iter, L=1, tol=tol, alpha=1.0)


def subset(p, idx):
    """Same problem data restricted to problems `idx` (array of indices)."""
    q = dict(p)
    q["x0"] = p["x0"][idx].copy()
    return q


def tassa_batch(B, N=150, dt=0.03, seed=_seed(6), I_o=50, I_a=5, L=40, tol=1e-3):
    """Tutorial problem (notebooks/Tutorial.ipynb cells 4-27): car parking of Tassa et al. with the pseudo-Huber cost
    (cell 14: cu = 1e-2*(1, .01), running cx = 1e-3*(1,1) with px = (.1,.1) on (x, y), final cf = (.1,.1,1,.3) with
    pf = (.01,.01,.01,1)), control limits |w| <= 0.5, |a| <= 2 (cell 25), rho_u = diag(1e-1, 1e-2), 5 ADMM
    iterations, 40 line-search candidates (cell 27).  The notebook uses N = 500 (15 s); the initial states here are
    its x0 = (1, 1, 3pi/2, 0) with a seeded perturbation per problem, the initial controls its N(0, 0.1^2) guess."""
    rng = np.random.default_rng(seed)
    n, m = 4, 2
    x0 = np.tile(np.array([1.0, 1.0, 1.5 * np.pi, 0.0]), (B, 1))
    x0[:, :2] += rng.uniform(-0.3, 0.3, (B, 2))
    x0[:, 2] += rng.uniform(-0.3, 0.3, B)
    u0 = rng.normal(0.0, 0.1, (N, m))
    zs = np.zeros((2, n))
    cxw = np.array([1e-3, 1e-3, 0.0, 0.0])
    seq = np.zeros(N, dtype=np.int32)
    seq[-1] = 1
    lo_u = np.tile(np.array([-0.5, -2.0]), (N, 1))
    return dict(name="tassa_car", model="tassa_car", dt=dt, N=N, n=n, m=m, cost="pseudo_huber", zs=zs, seq=seq,
                Qdiag=np.stack([cxw, cxw]), Hp=np.tile(np.array([0.1, 0.1, 1.0, 1.0]), (2, 1)),
                Qdiag_b=np.stack([np.zeros(n), np.array([0.1, 0.1, 1.0, 0.3])]),
                Hp_b=np.stack([np.ones(n), np.array([0.01, 0.01, 0.01, 1.0])]),
                Rdiag=1e-2 * np.array([1.0, 0.01]), u_std=1e-2, x0=x0, u0=u0,
                lo_u=lo_u, hi_u=-lo_u, lo_x=None, hi_x=None,
                rho_u=np.tile(np.array([1e-1, 1e-2]), (N, 1)), rho_x=None, I_o=I_o, I_a=I_a, L=L, tol=tol, alpha=1.0)


def parking_batch(B, N=500, dt=0.03, seed=_seed(7), I_o=10, I_a=10, L=50, tol=1e-1):
    """Parking between two parked cars: iLQR-ADMM with a STATE projection onto the outside of two rotated rectangles
    (notebooks/Car/Iterative LQR with state constraints.ipynb cells 4-20; README animation_state_bounds.gif).
    Problem 0 is the notebook's x0 = (0, -2, pi/2, 0)."""
    rng = np.random.default_rng(seed)
    n, m = 4, 2
    # start states: the notebook's, then four behind / beside the parked cars (their paths cross the obstacle sets, so
    # the inner projection ADMM really iterates), then seeded perturbations of those five
    anchors = np.array([[0.0, -2.0, np.pi / 2, 0.0], [-9.0, -1.0, -np.pi / 2, 0.0], [-1.0, -9.0, np.pi, 0.0],
                        [-10.0, -4.0, 0.0, 0.0], [-2.0, -10.0, np.pi / 2, 0.0]])
    x0 = anchors[np.arange(B) % 5].copy()
    if B > 5:
        x0[5:, :2] += rng.uniform(-0.4, 0.4, (B - 5, 2))
        x0[5:, 2] += rng.uniform(-0.3, 0.3, B - 5)
    zs = np.stack([np.zeros(n), np.array([-5.0, -5.0, np.pi / 4, 0.0])])
    Qdiag = np.stack([np.zeros(n), np.full(n, 1e2)])
    seq = np.zeros(N, dtype=np.int32)
    seq[-1] = 1
    a_safe = np.array([[2.0, 1.0], [2.0, 1.0]]) + 0.5                   # cell 18
    Ws = np.stack([np.diag(a_safe[i, 0] / a_safe[i]) for i in range(2)])
    al = -np.pi / 4
    Rm = np.array([[np.cos(al), -np.sin(al)], [np.sin(al), np.cos(al)]])
    Ws = Ws @ Rm.T
    rho_x = np.zeros((N, n))
    rho_x[:, :2] = 1e-1
    obstacles = dict(kind="square", centers=np.array([[-7.0, -3.0], [-3.0, -7.0]]), W=Ws, W_inv=np.linalg.inv(Ws),
                     lower=a_safe[:, 0] / 2, upper=1e5, rho=1e1, max_iter=15, threshold=1e-3)
    return dict(name="parking", model="car", dt=dt, N=N, n=n, m=m, zs=zs, Qdiag=Qdiag, seq=seq, u_std=1e-2,
                x0=x0, u0=np.zeros((N, m)), lo_u=None, hi_u=None, lo_x=None, hi_x=None, rho_u=None, rho_x=rho_x,
                obstacles=obstacles, I_o=I_o, I_a=I_a, L=L, tol=tol, alpha=1.0)


def arm_robust_batch(B, N=100, dt=0.01, seed=_seed(8), I_o=50, I_a=10, L=30, tol=1e-4, var_x0=0.1, prob=0.82):
    """Robust iSLS-ADMM on the planar 3-DoF arm (notebooks/3DoF robot/State bounds and robust control bounds.ipynb
    cells 12-26): reach final_pos = (1.5, 2, 0) with |u| <= 6 holding with probability `prob` under N(0, var_x0)
    perturbations of the initial joint positions; columns [d_u | Phi_u(:, :3)].  Problem 0 is the notebook's q0."""
    from scipy.stats import norm
    rng = np.random.default_rng(seed)
    n, m = 9, 3
    q0 = np.tile(np.array([np.pi / 3, -np.pi / 2, -np.pi / 4]), (B, 1))
    if B > 1:
        q0[1:] += rng.normal(0.0, 0.1, (B - 1, 3))
    a1 = q0[:, 0]
    a2 = a1 + q0[:, 1]
    a3 = a2 + q0[:, 2]
    fk = np.stack([np.cos(a1) + np.cos(a2) + np.cos(a3), np.sin(a1) + np.sin(a2) + np.sin(a3), np.zeros(B)], -1)
    x0 = np.concatenate([q0, np.zeros_like(q0), fk], axis=-1)
    zs = np.stack([np.zeros(n), np.array([0, 0, 0, 0, 0, 0, 1.5, 2.0, 0.0])])
    Qdiag = np.stack([np.zeros(n), np.array([0, 0, 0, 1e3, 1e3, 1e3, 1e3, 1e3, 0.0])])       # cell 12
    seq = np.zeros(N, dtype=np.int32)
    seq[-1] = 1
    dim = 3
    mu = np.zeros(1 + dim)
    mu[0] = 1.0
    sigma = np.zeros(1 + dim)
    sigma[1:] = var_x0
    psi_inv = norm.ppf(prob)
    Au = np.diag(np.sqrt(sigma))
    As = [np.concatenate([Au, (-mu / psi_inv)[None]], axis=0), np.concatenate([Au, (mu / psi_inv)[None]], axis=0)]
    bs = [np.append(np.zeros(1 + dim), 6.0 / psi_inv), np.append(np.zeros(1 + dim), 6.0 / psi_inv)]
    robust = dict(dim=dim, As=As, bs=bs, rho_u=1.0, inner_rho=1e1, inner_max_iter=100, inner_threshold=1e-4)
    return dict(name="arm3_robust", model="arm3", dt=dt, N=N, n=n, m=m, zs=zs, Qdiag=Qdiag, seq=seq, u_std=1e-4,
                x0=x0, u0=np.zeros((N, m)), lo_u=None, hi_u=None, lo_x=None, hi_x=None, rho_u=None, rho_x=None,
                robust=robust, I_o=I_o, I_a=I_a, L=L, tol=tol, alpha=1.0)


def arm_robust_x_batch(B, comps=(3, 4), v_max=(0.9, 0.8), rho_x=10.0, project_u=True, **kw):
    """arm_robust_batch plus robust STATE bounds through isls_admm's project_x (isls/isls.py:631-638): the joint
    velocities listed in `comps` must stay within +-v_max at the same chance level, i.e. the rows
    [x_nom + d_x | Phi_x(:, :3)] of those components are projected onto the same two cones as the controls with their
    own offsets; Qr = diag(rho_x) on those components, 0 elsewhere."""
    from scipy.stats import norm
    p = arm_robust_batch(B, **kw)
    rb = p["robust"]
    psi_inv = norm.ppf(0.82 if "prob" not in kw else kw["prob"])
    dim = rb["dim"]
    bs = [[np.append(np.zeros(1 + dim), v / psi_inv), np.append(np.zeros(1 + dim), v / psi_inv)] for v in v_max]
    rx = np.zeros((p["N"], p["n"]))
    rx[:, list(comps)] = rho_x
    rb["x"] = dict(comps=list(comps), bs=np.array(bs), rho_x=rx)
    rb["u_unprojected"] = not project_u
    p["name"] = "arm3_robust_x"
    return p


def di_obstacle_batch(B=1, N=100, seed=_seed(9), max_iter=200, tol=1e-3):
    """LQT-ADMM double integrator (n=4, m=2) avoiding two spherical obstacles with a STATE projection
    (notebooks/Double integrator/LQR and SLS with spherical obstacle avoidance.ipynb cells 4-14, scenario 0):
    project_set_convex (5 iterations, threshold 1e-2) followed by Dykstra (50 iterations, tol 1e-5) over the quadratic
    shells 0.5 ||p - c_k||^2 >= 0.5 (1.1 r_k)^2 of the position.  Problem 0 is the notebook's (x0 = 0, target (1,1))."""
    rng = np.random.default_rng(seed)
    n, m = 4, 2
    dt = 1.0 / N
    zs = np.stack([np.zeros(n), np.array([1.0, 1.0, 0.0, 0.0])])
    Qdiag = np.stack([np.zeros(n), np.full(n, 1e3)])
    seq = np.zeros(N, dtype=np.int32)
    seq[-1] = 1
    x0 = np.zeros((B, n))
    if B > 1:
        x0[1:, :2] = rng.uniform(-0.1, 0.3, (B - 1, 2))
    rho_x = np.zeros((N, n))
    rho_x[:, :2] = 1.0
    radii = np.array([0.1, 0.15]) * 1.1
    obstacles = dict(kind="quadratic", centers=np.array([[0.5, 0.5], [0.5, 0.2]]), lower=0.5 * radii ** 2, upper=1e2,
                     rho=1.0, max_iter=5, threshold=1e-2, dykstra_max_iter=50, dykstra_tol=1e-5)
    return dict(name="di_obstacles", model="double_integrator", dt=dt, N=N, n=n, m=m, zs=zs, Qdiag=Qdiag, seq=seq,
                u_std=1e-4, x0=x0, u0=np.zeros((N, m)), lo_u=None, hi_u=None, lo_x=None, hi_x=None, rho_u=None,
                rho_x=rho_x, obstacles=obstacles, I_o=1, I_a=max_iter, L=1, tol=tol, alpha=1.0)


def sls_state_bounds_problem(N=100):
    """SLS-ADMM with robust control bounds AND a robust terminal state constraint (notebooks/Double integrator/LQR
    and SLS with state bounds.ipynb cells 3-17): 1-D double integrator (n=2, m=1), no tracking cost (Q = 0: the
    final state is imposed by the constraints only), |u| <= 3, final position in [0.5, 0.5], final velocity 0, all as
    chance constraints (p = 0.9, var_x0 = 0.02) on the rows of [d | Phi(:, :1)]."""
    from scipy.stats import norm
    n, m = 2, 1
    zs = np.stack([np.zeros(n), np.array([1.0, 1.0])])
    seq = np.zeros(N, dtype=np.int32)
    seq[-1] = 1
    mu = np.zeros(2)
    mu[0] = 1.0
    sigma = np.zeros(2)
    sigma[1:] = 0.02
    psi = norm.ppf(0.9)
    Au = np.diag(np.sqrt(sigma))
    As = [np.concatenate([Au, (-mu / psi)[None]], 0), np.concatenate([Au, (mu / psi)[None]], 0)]
    cone = lambda up, lo: [np.append(np.zeros(2), up / psi), np.append(np.zeros(2), -lo / psi)]
    rho_x = np.zeros(N * n)
    rho_x[-2:] = 1e3
    return dict(n=n, m=m, N=N, dt=1.0 / N, zs=zs, Qdiag=np.zeros((2, n)), seq=seq, u_std=1e-4, As=As,
                bs_u=cone(3.0, -3.0), x_rows=[(N * n - 2, As, cone(0.5, 0.5)), (N * n - 1, As, cone(0.0, 0.0))],
                rho_x=rho_x, rho_u=1e-3, max_iter=100, tol=1e-5, inner_rho=1e1, inner_max_iter=20, inner_threshold=1e-2)
