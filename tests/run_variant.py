#!/usr/bin/env python
"""Helper of tests/test_gpu_kernel_variants.py: solves two small problems through the public API in THIS process (whose
environment selects the kernel variants: ISLS_FF_STAGES, ISLS_ADMM_STAGES are read once per process by the library)
and saves every output array."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ilqr-admm_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import gpu_util as G
from isls_b200 import configs

out = {}
for tag, p in (("arm", configs.arm_batch(40, I_o=4, I_a=4)),                      # state + control bounds: k_admm
               ("park", configs.parking_batch(6, N=120, dt=0.125, I_o=3, I_a=4)),     # obstacle sets: k_admm + k_obst_project
               ("car", configs.car_batch(70, I_o=3, I_a=3))):                       # control bounds only (fused update)
    o = G.run_ilqr_admm(p)
    for k, v in o.items():
        out[tag + "_" + k] = v
o = G.run_isls_admm(configs.arm_robust_batch(3, I_o=3, I_a=3))      # robust iSLS-ADMM: k_isls_cols
for k, v in o.items():
    out["robust_" + k] = v
o = G.run_lqt_admm_dp(configs.di_batch(5, max_iter=60))                 # LQT-ADMM: k_lqt_admm<., SM>
for k, v in o.items():
    out["lqt_" + k] = v
o = G.run_lqt_admm_dp(configs.di_obstacle_batch(2, max_iter=12))        # ... with the spherical-obstacle projection
for k, v in o.items():
    out["lqtobs_" + k] = v
np.savez(sys.argv[1], **out)
print("variant ok", sorted(k for k in os.environ if k.startswith("ISLS_")))
