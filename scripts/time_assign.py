import sys, torch
sys.path.insert(0, ".")
from ark_analysis_b200 import som as S
n, C, K = 50 * 1024 * 1024, 32, 100
X = torch.rand((n, C), device="cuda")
W = X[torch.randperm(n, device="cuda")[:K]].contiguous()
lab = torch.empty(n, dtype=torch.int32, device="cuda")
for _ in range(3):
    S.bmu(X, W, labels=lab)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    S.bmu(X, W, labels=lab)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"assign {ms:.3f} ms  {n*132/ms/1e6/6548.2:.3f} of roofline")
