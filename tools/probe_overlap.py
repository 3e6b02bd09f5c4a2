#!/usr/bin/env python
"""One slot of the overlapped schedule under the stopwatch (isls_probe_overlap_f64): C5 car, 65,536 problems."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ilqr-admm_b200")):
    sys.path.insert(0, p)
from isls_b200 import configs, solver as S

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
p = configs.car_batch(B, tol=1e-3, I_o=20, I_a=5, L=20)
plan = S.Plan("car", p["N"], 4, 2, p["dt"], p["Qdiag"], p["seq"], p["u_std"], p["L"], rho_u=p["rho_u"], lo_u=p["lo_u"],
              hi_u=p["hi_u"])
sv = S.BatchSolver(plan, B, "cuda:0", max_outer=20, max_admm=5, logs=False)
sv.set_inputs(p["x0"], p["u0"], p["zs"])
for ls in (1, 2, 3):
    for ff in (2, 3, 4, 6):
        r = sv.probe_overlap(ls, ff)
        print(json.dumps(dict(ls_ctas=ls, ff_depth=ff, **{k: round(v, 4) for k, v in r.items()})), flush=True)
