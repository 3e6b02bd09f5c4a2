#!/usr/bin/env python
"""Stage-level generic-operator Riccati pass (isls_riccati_f64), arm shape n=9, m=3, N=100, 16,384 problems - the
command profiled with ncu for k_riccati_generic."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ilqr-admm_b200")):
    sys.path.insert(0, p)
import torch
from isls_b200 import solver as S

n_, m_, Bq, N_ = 9, 3, 16384, 100
g = torch.Generator(device="cuda").manual_seed(1)
A_ = torch.eye(n_, dtype=torch.float64, device="cuda").expand(Bq, N_, n_, n_).contiguous()
A_ += 0.05 * torch.randn(Bq, N_, n_, n_, dtype=torch.float64, device="cuda", generator=g)
B_ = 0.1 * torch.randn(Bq, N_, n_, m_, dtype=torch.float64, device="cuda", generator=g)
c_ = torch.randn(Bq, N_, n_ + m_, dtype=torch.float64, device="cuda", generator=g)
W_ = torch.randn(Bq, N_, n_ + m_, n_ + m_, dtype=torch.float64, device="cuda", generator=g)
C_ = W_ @ W_.transpose(-1, -2) + torch.eye(n_ + m_, dtype=torch.float64, device="cuda")
for _ in range(3):
    K, k, bad = S.riccati(A_, B_, c_, C_)
torch.cuda.synchronize()
print("riccati generic ok", float(K.abs().max()))
