#!/usr/bin/env python
"""C1 (LQT-ADMM double integrator, 32 problems) - the command profiled with ncu for k_lqt_admm."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ilqr-admm_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import gpu_util as G
from isls_b200 import configs

o = G.run_lqt_admm_dp(configs.di_batch(32, max_iter=300))
torch.cuda.synchronize()
print("C1 small ok", int(o["admm_iters"].max()))
