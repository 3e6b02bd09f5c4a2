"""TEST INFRASTRUCTURE (oracle) - numpy statements of the three dynamics models of the hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this.

Every function is vectorised over arbitrary leading axes (``x[..., n]``, ``u[..., m]``) so the same code serves
the single-problem reference shim (leading axis = line-search candidates or time steps) and the batched oracle
(leading axes = problems x candidates).

Formulas follow the reference's notebooks (the reference keeps its models in notebooks, not in the package):
  * car   : notebooks/Car/Iterative LQR with control constraints.ipynb cell 6
  * arm3  : notebooks/3DoF robot/State and control bound constraints.ipynb cells 9-10 with the pinocchio
            FK / LOCAL_WORLD_ALIGNED Jacobian replaced by the closed form of the planar 3R chain described by
            notebooks/3DoF robot/urdfs/3dof_robot.urdf:73-102 (three unit links about z)
  * double_integrator : isls/utils.py:266-276 (get_double_integrator_AB) + isls/sls_base.py:49-53
"""
from math import factorial

import numpy as np

TWO_PI = 2.0 * np.pi


# ----------------------------------------------------------------------------------------------- double integrator
def double_integrator_AB(nb_dim, nb_deriv=2, dt=0.01):
    """A = kron(A1d, I), B = kron(B1d, I) with the Taylor blocks of utils.py:266-276."""
    A1 = np.zeros((nb_deriv, nb_deriv))
    for i in range(nb_deriv):
        A1 += np.diag(np.ones(nb_deriv - i), i) * dt ** i / factorial(i)
    B1 = np.zeros((nb_deriv, 1))
    for i in range(1, nb_deriv + 1):
        B1[nb_deriv - i] = dt ** i / factorial(i)
    return np.kron(A1, np.eye(nb_dim)), np.kron(B1, np.eye(nb_dim))


class DoubleIntegrator:
    name = "double_integrator"

    def __init__(self, nb_dim=2, dt=0.02):
        self.nb_dim, self.dt = nb_dim, dt
        self.n, self.m = 2 * nb_dim, nb_dim
        self.A, self.B = double_integrator_AB(nb_dim, 2, dt)

    def f(self, x, u):
        return x @ self.A.T + u @ self.B.T

    def get_AB(self, x, u):
        lead = x.shape[:-1]
        A = np.broadcast_to(self.A, lead + self.A.shape).copy()
        B = np.broadcast_to(self.B, lead + self.B.shape).copy()
        return A, B


# ------------------------------------------------------------------------------------------------------------ car
class Car:
    """Simple kinematic car, state [x, y, theta, v], control [steer-rate-like, accel]."""
    name = "car"
    n, m = 4, 2

    def __init__(self, dt=0.1):
        self.dt = dt

    def f(self, x, u):
        dt = self.dt
        px = x[..., 0] + dt * x[..., 3] * np.cos(x[..., 2])
        py = x[..., 1] + dt * x[..., 3] * np.sin(x[..., 2])
        th = x[..., 2] + dt * x[..., 3] * u[..., 0]
        v = x[..., 3] + dt * u[..., 1]
        th = np.mod(th, TWO_PI)          # keep theta in [0, 2 pi) as the notebook does (x3 % (2*np.pi))
        return np.stack([px, py, th, v], axis=-1)

    def get_AB(self, x, u):
        dt = self.dt
        lead = x.shape[:-1]
        A = np.zeros(lead + (4, 4))
        B = np.zeros(lead + (4, 2))
        for i in range(4):
            A[..., i, i] = 1.0
        s, c = np.sin(x[..., 2]), np.cos(x[..., 2])
        A[..., 0, 2] = dt * x[..., 3] * -s
        A[..., 1, 2] = dt * x[..., 3] * c
        A[..., 0, 3] = dt * c
        A[..., 1, 3] = dt * s
        A[..., 2, 3] = dt * u[..., 0]
        B[..., 2, 0] = dt * x[..., 3]
        B[..., 3, 1] = dt
        return A, B


# ----------------------------------------------------------------------------------------------------------- arm3
def fk3(q):
    """End-effector position of the planar 3R chain with unit links: (sum cos(cumsum q), sum sin(cumsum q), 0)."""
    a1 = q[..., 0]
    a2 = a1 + q[..., 1]
    a3 = a2 + q[..., 2]
    px = np.cos(a1) + np.cos(a2) + np.cos(a3)
    py = np.sin(a1) + np.sin(a2) + np.sin(a3)
    return np.stack([px, py, np.zeros_like(px)], axis=-1)


def jac3(q):
    """d fk3 / d q, shape [..., 3, 3] (last row zero)."""
    a1 = q[..., 0]
    a2 = a1 + q[..., 1]
    a3 = a2 + q[..., 2]
    s1, s2, s3 = np.sin(a1), np.sin(a2), np.sin(a3)
    c1, c2, c3 = np.cos(a1), np.cos(a2), np.cos(a3)
    J = np.zeros(q.shape[:-1] + (3, 3))
    J[..., 0, 0] = -(s1 + s2 + s3)
    J[..., 0, 1] = -(s2 + s3)
    J[..., 0, 2] = -s3
    J[..., 1, 0] = c1 + c2 + c3
    J[..., 1, 1] = c2 + c3
    J[..., 1, 2] = c3
    return J


class Arm3:
    """Planar 3-DoF arm, state [q(3), qdot(3), p_ee(3)], control qddot(3)."""
    name = "arm3"
    n, m = 9, 3

    def __init__(self, dt=0.01):
        self.dt = dt
        self.A6, self.B6 = double_integrator_AB(3, 2, dt)

    def f(self, x, u):
        dt = self.dt
        q = x[..., 0:3] + x[..., 3:6] * dt + 0.5 * u * (dt ** 2)
        qd = x[..., 3:6] + u * dt
        return np.concatenate([q, qd, fk3(q)], axis=-1)

    def get_AB(self, x, u):
        dt = self.dt
        lead = x.shape[:-1]
        A = np.zeros(lead + (9, 9))
        B = np.zeros(lead + (9, 3))
        A[..., :6, :6] = self.A6
        B[..., :6, :] = self.B6
        J = jac3(x[..., 0:3] + x[..., 3:6] * dt + 0.5 * u * (dt ** 2))
        A[..., 6:, 0:3] = J
        A[..., 6:, 3:6] = J * dt
        B[..., 6:, :] = 0.5 * J * (dt ** 2)
        return A, B

    def state_from_q(self, q0):
        q0 = np.asarray(q0, dtype=np.float64)
        return np.concatenate([q0, np.zeros_like(q0), fk3(q0)], axis=-1)


def make_model(name, **kw):
    if name == "double_integrator":
        return DoubleIntegrator(**kw)
    if name == "car":
        return Car(**kw)
    if name == "arm3":
        return Arm3(**kw)
    if name == "tassa_car":
        return TassaCar(**kw)
    raise KeyError(name)


# ------------------------------------------------------------------------------------------------- Tassa parking car
class TassaCar:
    """Car-parking model of Tassa et al. used by the reference's Tutorial (notebooks/Tutorial.ipynb cell 8):
    state [x, y, car angle, front-wheel velocity], control [front-wheel angle, acceleration], axle distance 2.0.
    The notebook differentiates it with autograd; get_AB below is the hand-derived Jacobian."""
    name = "tassa_car"
    n, m = 4, 2
    dist = 2.0

    def __init__(self, dt=0.03):
        self.dt = dt

    def f(self, s, u):
        dt, d = self.dt, self.dist
        w, a = u[..., 0], u[..., 1]
        x, y, o, v = s[..., 0], s[..., 1], s[..., 2], s[..., 3]
        f_ = dt * v
        ins = d ** 2 - (np.sin(w) * f_) ** 2
        b = f_ * np.cos(w) + d - np.sqrt(ins)
        do = np.arcsin(np.sin(w) * f_ / d)
        return np.stack([x + b * np.cos(o), y + b * np.sin(o), o + do, v + a * dt], axis=-1)

    def get_AB(self, s, u):
        dt, d = self.dt, self.dist
        w = u[..., 0]
        o, v = s[..., 2], s[..., 3]
        f_ = dt * v
        sw, cw = np.sin(w), np.cos(w)
        S = np.sqrt(d ** 2 - (sw * f_) ** 2)
        b = f_ * cw + d - S
        db_dv = dt * (cw + sw * sw * f_ / S)
        db_dw = -f_ * sw + sw * cw * f_ * f_ / S
        lead = s.shape[:-1]
        A = np.zeros(lead + (4, 4))
        B = np.zeros(lead + (4, 2))
        for i in range(4):
            A[..., i, i] = 1.0
        so, co = np.sin(o), np.cos(o)
        A[..., 0, 2] = -b * so
        A[..., 1, 2] = b * co
        A[..., 0, 3] = db_dv * co
        A[..., 1, 3] = db_dv * so
        A[..., 2, 3] = dt * sw / S
        B[..., 0, 0] = db_dw * co
        B[..., 1, 0] = db_dw * so
        B[..., 2, 0] = f_ * cw / S
        B[..., 3, 1] = dt
        return A, B
