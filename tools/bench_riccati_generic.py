import sys
sys.path[:0]=["/root/repo","/root/repo/ilqr-admm_b200"]
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
K, k, bad = S.riccati(A_, B_, c_, C_); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): K, k, bad = S.riccati(A_, B_, c_, C_)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
byts = Bq * N_ * 8.0 * (n_ * n_ + n_ * m_ + (n_ + m_) + (n_ + m_) ** 2 + m_ * n_ + m_)
print("riccati generic (9,3) ms %.3f  GB/s %.1f frac %.3f" % (ms, byts / ms / 1e6, byts / ms / 1e6 / 6542.7), "checksum %.17g %.17g" % (float(K.double().sum()), float(k.double().sum())), int(bad.sum()))
