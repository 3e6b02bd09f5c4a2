#!/usr/bin/env python
"""Car iLQR-ADMM (C5 shard size: 8,192 problems by default) with a short fixed budget - the command profiled with ncu
for the small-batch kernels.  python tools/run_car_small.py [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ilqr-admm_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import gpu_util as G
from isls_b200 import configs

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
p = configs.car_batch(B, I_o=2, I_a=3, L=20)
o = G.run_ilqr_admm(p, fixed_budget=True, want_masks=False)
torch.cuda.synchronize()
print("car small ok", B, float(o["cost"].mean()))
