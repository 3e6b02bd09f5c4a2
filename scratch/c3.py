import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT,'ilqr-admm_b200'), os.path.join(ROOT,'tools')): sys.path.insert(0,p)
import torch
from isls_b200 import configs, solver as S
import bench_configs as BC
p=configs.arm_batch(16384, I_o=3, I_a=3)
ms,out,prof=BC.isls_case(p, True)
print(ms, prof)
