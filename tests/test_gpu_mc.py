"""GPU: Monte-Carlo closed-loop evaluation - get_trajectory_batch / dp / sls on the device against
the reference goldens (noise-free), the oracle at 10,000 samples, and the statistics of the process noise."""
import numpy as np
import pytest

from oracle import models as M, restated as R

pytestmark = pytest.mark.gpu


def _np(t):
    return t.cpu().numpy()


def test_mc_rollouts_vs_reference_golden(golden):
    from isls_b200 import SLS, get_double_integrator_AB, iSLS
    g = golden("mc_rollouts")
    s = SLS(4, 2, 30)
    s.AB = get_double_integrator_AB(2, 2, 0.05)
    for mode, args in (("dp", (g["di_K"], g["di_k"])), ("sls", (g["di_Ks"], g["di_ks"])), ("batch", (g["di_us"],))):
        x, u = getattr(s, "get_trajectory_" + mode)(g["di_x0"], *args)
        assert np.abs(_np(x) - g["di_%s_x" % mode]).max() < 1e-12
        assert np.abs(_np(u) - g["di_%s_u" % mode]).max() < 1e-12
    c = iSLS(4, 2, 25)
    c.forward_model = ("car", {"dt": 0.1})
    import torch
    c.x_nom, c.u_nom = torch.as_tensor(g["car_x_nom"]), torch.as_tensor(g["car_u_nom"])
    x, u = c.get_trajectory_dp(g["car_x0"], g["car_K"], g["car_k"])
    assert np.abs(_np(x) - g["car_dp_x"]).max() < 1e-11 and np.abs(_np(u) - g["car_dp_u"]).max() < 1e-11
    x, u = c.get_trajectory_sls(g["car_x0"], g["car_Ks"], g["car_ks"])
    assert np.abs(_np(x) - g["car_sls_x"]).max() < 1e-11 and np.abs(_np(u) - g["car_sls_u"]).max() < 1e-11
    # single initial state -> unbatched result, like the reference
    x1, u1 = s.get_trajectory_dp(g["di_x0"][0], g["di_K"], g["di_k"])
    assert x1.shape == (30, 4) and np.abs(_np(x1) - g["di_dp_x"][0]).max() < 1e-12


def test_mc_10000_samples_vs_oracle_and_noise_statistics():
    """The notebooks' robustness experiment shape: 10,000 sampled initial positions through one SLS controller
    (Double integrator/LQR and SLS with control bounds.ipynb cells 19-20); then the process-noise generator."""
    from isls_b200 import SLS, get_double_integrator_AB
    rng = np.random.default_rng(3)
    n, m, N, dt = 4, 2, 50, 0.02
    A, B = M.double_integrator_AB(2, 2, dt)
    Ks = (np.tril(np.ones((N, N)))[:, None, :, None] * rng.normal(0, 0.05, (N, m, N, n))).reshape(N * m, N * n)
    ks = rng.normal(0, 0.5, N * m)
    x0 = np.zeros((10000, n)); x0[:, :2] = rng.normal(0, 0.1, (10000, 2))
    s = SLS(n, m, N)
    s.AB = get_double_integrator_AB(2, 2, dt)
    x, u = s.get_trajectory_sls(x0, Ks, ks)
    xo, uo = R.mc_rollout(M.make_model("double_integrator", nb_dim=2, dt=dt), "sls", x0, Ks, ks, N)
    assert np.abs(_np(u) - uo).max() / np.abs(uo).max() < 1e-12 and np.abs(_np(x) - xo).max() / np.abs(xo).max() < 1e-12
    # noise: x_{t+1} = A x_t + B u_t + w, w ~ N(0, 0.05) i.i.d.: recover w and test its moments / independence
    K0, k0 = np.zeros((N, m, n)), np.zeros((N, m))
    xs, us = s.get_trajectory_dp(np.zeros((20000, n)), K0, k0, noise_scale=0.05, seed=7)
    xs = _np(xs)
    w = xs[:, 1:] - xs[:, :-1] @ A.T
    assert abs(w.mean()) < 2e-4 and abs(w.std() - 0.05) < 2e-4
    flat = w.reshape(20000, -1)
    cc = np.corrcoef(flat[:, :16].T)
    assert np.abs(cc - np.eye(16)).max() < 0.04
    kurt = np.mean((w / 0.05) ** 4)
    assert abs(kurt - 3.0) < 0.05
    xs2, _ = s.get_trajectory_dp(np.zeros((20000, n)), K0, k0, noise_scale=0.05, seed=7)
    assert np.array_equal(xs, _np(xs2))                       # reproducible from the seed
    xs3, _ = s.get_trajectory_dp(np.zeros((20000, n)), K0, k0, noise_scale=0.05, seed=8)
    assert not np.array_equal(xs, _np(xs3))
