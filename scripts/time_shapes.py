"""Training-pass and assign timings for the BASELINE shapes that do not fit the whole-pass kernel
(cfg3: C=40, K=400; cfg4: C=100, K=100) beside cfg2, Pixie-like rows, CUDA events."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import bench  # noqa: E402
from ark_analysis_b200 import som as S  # noqa: E402
from conftest import pixie_like  # noqa: E402

peak = bench.measured_peaks()[0]


def timed(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for name, n_train, n_assign, C, xd in (("cfg2", 5241600, 52428800, 32, 10),
                                        ("cfg3 (8 FOVs)", 3355392, 33554432, 40, 20),
                                        ("cfg4", 5000064, 5000064, 100, 10)):
    K = xd * xd
    base = torch.from_numpy(pixie_like(1 << 20, C)).cuda()
    reps_a = (n_assign + base.shape[0] - 1) // base.shape[0]
    X = base.repeat(reps_a, 1)[:n_assign].contiguous()
    Xt = X[:n_train]
    idx = np.random.default_rng(42).choice(1 << 20, K, replace=False)
    W0 = base[torch.from_numpy(idx).cuda()].double()
    W = S.train_som(Xt, W0, xd, xd, rlen=1, batches_per_pass=32)
    ms_t = timed(lambda: S.train_som(Xt, W0, xd, xd, rlen=1, batches_per_pass=32), 5)
    W32 = W.float().contiguous()
    lab = torch.empty(n_assign, dtype=torch.int32, device="cuda")
    ms_a = timed(lambda: S.bmu(X, W32, labels=lab), 5)
    gbs = n_assign * (4 * C + 4) / ms_a / 1e6
    print(f"{name}: C={C} K={K} train pass ({n_train} rows, 32 steps) {ms_t:.3f} ms = "
          f"{n_train/ms_t/1e6:.2f} Gpx/s; assign ({n_assign} rows) {ms_a:.3f} ms = "
          f"{n_assign/ms_a/1e6:.2f} Gpx/s, {gbs:.0f} GB/s, frac={gbs/peak:.3f}", flush=True)
    del X, Xt, lab
    torch.cuda.empty_cache()
