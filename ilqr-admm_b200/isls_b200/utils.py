"""Helpers of the reference that belong to the hot path's API surface (isls/utils.py)."""
from math import factorial

import numpy as np


def get_double_integrator_AB(nb_dim, nb_deriv=2, dt=0.01):
    """isls/utils.py:266-276."""
    A1 = np.zeros((nb_deriv, nb_deriv))
    for i in range(nb_deriv):
        A1 += np.diag(np.ones(nb_deriv - i), i) * dt ** i / factorial(i)
    B1 = np.zeros((nb_deriv, 1))
    for i in range(1, nb_deriv + 1):
        B1[nb_deriv - i] = dt ** i / factorial(i)
    return np.kron(A1, np.eye(nb_dim)), np.kron(B1, np.eye(nb_dim))


def find_mus(zs, seq):
    """isls/utils.py:95-99."""
    return np.stack([zs[s] for s in seq]).flatten()


def diag_of(M, name):
    """Diagonal of a (stack of) square matrices; raises if any off-diagonal entry is non-zero (the device path
    carries diagonal weights only)."""
    M = np.asarray(M, dtype=np.float64)
    d = np.diagonal(M, axis1=-2, axis2=-1)
    off = M - np.einsum("...i,ij->...ij", d, np.eye(M.shape[-1]))
    if np.any(off != 0.0):
        raise NotImplementedError("%s must be diagonal on the device path" % name)
    return np.ascontiguousarray(d)
