#!/usr/bin/env python
"""Diagnostic for tests/test_gpu_kernel_variants.py: runs tests/run_variant.py once per kernel-variant environment and
prints, per variant, the output arrays that are not bit-identical to the plain kernels (max |diff|, count)."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VARIANTS = {
    "plain": {"ISLS_FF_STAGES": "0", "ISLS_ADMM_STAGES": "0", "ISLS_COLS_STAGES": "0", "ISLS_LQT_SMEM": "0", "ISLS_OVERLAP": "0", "ISLS_SLS_CTRL_DENSE": "1", "ISLS_ADMM_LOOP": "0"},
    "auto": {},
    "tma_nojc": {"ISLS_FF_MODE": "2", "ISLS_FF_JC": "0"},
    "staged": {"ISLS_FF_MODE": "0"},
    "overlap": {"ISLS_OVERLAP": "1"},
    "loop": {"ISLS_ADMM_LOOP": "2"},
    "loop_split": {"ISLS_ADMM_LOOP": "2", "ISLS_ADMM_LOOP_SPLIT": "1"},
}
td = tempfile.mkdtemp()
res = {}
for name, extra in VARIANTS.items():
    env = {k: v for k, v in os.environ.items() if not k.startswith("ISLS_")}
    env.update(extra)
    out = os.path.join(td, name + ".npz")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "run_variant.py"), out], env=env, capture_output=True,
                       text=True)
    if r.returncode:
        print(name, "FAILED", r.stdout[-2000:], r.stderr[-2000:])
        continue
    res[name] = np.load(out)
base = res["plain"]
for name, z in res.items():
    if name == "plain":
        continue
    bad = []
    for k in base.files:
        a, b = base[k], z[k]
        if not np.array_equal(a, b, equal_nan=True):
            fa, fb = a.astype(np.float64), b.astype(np.float64)
            m = ~(np.isnan(fa) & np.isnan(fb))
            d = np.abs(np.where(m, fa - fb, 0.0))
            bad.append("%s max|d|=%.3e n=%d" % (k, np.nanmax(d), int((d > 0).sum() + (np.isnan(fa) != np.isnan(fb)).sum())))
    print(name, "identical to plain" if not bad else "DIFFERS: " + "; ".join(bad))
