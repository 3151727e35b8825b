"""Warm timings of the pieces of one training step (CUDA events, 200 repetitions each)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from ark_analysis_b200 import som as S
sys.path.insert(0, "tests")
from conftest import pixie_like

C = int(sys.argv[1]) if len(sys.argv) > 1 else 32
xd = int(sys.argv[2]) if len(sys.argv) > 2 else 10
n = int(sys.argv[3]) if len(sys.argv) > 3 else 5241600
K = xd * xd
base = torch.from_numpy(pixie_like(1 << 20, C)).cuda()
X = base.repeat((n + (1 << 20) - 1) >> 20, 1)[:n].contiguous()
W64 = X[:K].to(torch.float64).clone()
W32 = X[:K].clone()
SN = torch.zeros((K, C + 1), dtype=torch.float64, device="cuda")
W = S.train_som(X, W64, xd, xd, rlen=1, batches_per_pass=32)
W32 = W.to(torch.float32)


def timeit(fn, reps=200):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


print("accum (memset+prep+bmu_tc fused sums+fixup) us:", timeit(lambda: S.som_accum(X, W32, 3, 32, SN=SN)))
print("apply us:", timeit(lambda: S.som_apply(W64, W32, SN, xd, xd, 1.5, 0.03)))
lab = torch.empty(n // 32 + 128, dtype=torch.int32, device="cuda")
Xs = X[: n // 32]
print("bmu only on n/32 contiguous rows us:", timeit(lambda: S.bmu(Xs, W32, labels=lab[: Xs.shape[0]])))
print("bmu+sums on n/32 contiguous rows us:", timeit(lambda: S.bmu(Xs, W32, labels=lab[: Xs.shape[0]], want_sums=True)))
print("full train pass ms:", timeit(lambda: S.train_som(X, W64, xd, xd, rlen=1, batches_per_pass=32), 20) / 1e3)
e = torch.empty(1, device="cuda")
print("torch tiny kernel us:", timeit(lambda: e.zero_()))
