"""GPU: the small-batch kernels (TMA-staged k_ff_tma with the Jacobian cache, cp.async-staged k_ff_staged, k_admm_staged,
k_isls_cols_staged - selected automatically below 1,536 tiles) against the plain kernels the large batches use (k_ff, k_admm,
k_isls_cols).  The variants differ in how operands reach the
registers, not in arithmetic: every output must agree BIT FOR BIT.  The library reads its variant switches once per
process, so each variant runs in its own subprocess (tests/run_variant.py)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _run(tmp_path, name, env_extra):
    env = {k: v for k, v in os.environ.items() if not k.startswith("ISLS_")}
    env.update(env_extra)
    out = str(tmp_path / (name + ".npz"))
    r = subprocess.run([sys.executable, os.path.join(HERE, "run_variant.py"), out], env=env, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    return np.load(out)


def test_staged_and_plain_kernels_agree_bitwise(tmp_path):
    auto = _run(tmp_path, "auto", {})
    plain = _run(tmp_path, "plain", {"ISLS_FF_STAGES": "0", "ISLS_ADMM_STAGES": "0", "ISLS_COLS_STAGES": "0",
                                   "ISLS_LQT_SMEM": "0", "ISLS_OVERLAP": "0", "ISLS_SLS_CTRL_DENSE": "1"})
    deep2 = _run(tmp_path, "ff2", {"ISLS_FF_STAGES": "2"})
    tma_nojc = _run(tmp_path, "tma_nojc", {"ISLS_FF_MODE": "2", "ISLS_FF_JC": "0"})    # k_ff_tma recomputing the Jacobian
    staged = _run(tmp_path, "staged", {"ISLS_FF_MODE": "0"})                           # round-1 rule: cp.async staging
    ovl = _run(tmp_path, "overlap", {"ISLS_OVERLAP": "1"})        # two-stream overlapped schedule forced at this size
    loop = _run(tmp_path, "loop", {"ISLS_ADMM_LOOP": "2"})        # whole inner ADMM loop in one persistent launch (k_admm_loop)
    assert set(auto.files) == set(plain.files)
    for k in auto.files:
        for other, nm in ((plain, "plain"), (deep2, "ff depth 2"), (tma_nojc, "TMA ff without the Jacobian cache"),
                          (staged, "cp.async-staged ff"), (ovl, "overlapped schedule"),
                          (loop, "inner loop in one launch")):
            assert np.array_equal(auto[k], other[k], equal_nan=True), "%s differs between auto and %s" % (k, nm)
    assert auto["arm_n_log"].min() >= 1 and np.isfinite(auto["park_cost"]).all()
