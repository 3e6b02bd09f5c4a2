#!/usr/bin/env python
"""C3 (arm iLQR-ADMM, 16,384 problems) with a short fixed budget - the command profiled with ncu for the
small-batch (latency-bound) kernels: k_ff_staged<Arm3Model,3>, k_admm<Arm3Model>."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ilqr-admm_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import gpu_util as G
from isls_b200 import configs

p = configs.arm_batch(16384, I_o=2, I_a=3, L=5)
o = G.run_ilqr_admm(p, fixed_budget=True, want_masks=False)
torch.cuda.synchronize()
print("C3 small ok", float(o["cost"].mean()))
