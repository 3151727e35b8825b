"""Assign kernel variants: parity, candidate-window margin and timing for one configuration
(environment knobs PIXIE_TC_STAGES / PIXIE_DELTA_SCALE are read by the library).
Usage: python scripts/variant_experiment.py <parity|timing|timingU|margin> [C] [K] [nfov]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402  (data generator of the benchmark)
from ark_analysis_b200 import som as S  # noqa: E402

cfg = {k: os.environ.get(k, "-") for k in
       ("PIXIE_TC_STAGES", "PIXIE_DELTA_SCALE")}
tag = " ".join(f"{k[6:]}={v}" for k, v in cfg.items())


def pixie_rows(nfov, C):
    bench.C = C
    n = nfov * bench.HW * bench.HW
    X = torch.empty((n, C), device="cuda", dtype=torch.float32)
    bench.gen_fovs_device(torch, "cuda", list(range(nfov)), X)
    return X


def trained_codebook(X, K):
    xd = int(round(np.sqrt(K)))
    n = min(X.shape[0], 1 << 20)
    idx = np.random.default_rng(42).choice(n, K, replace=False)
    W0 = X[torch.from_numpy(idx).cuda()].double().cpu().numpy()
    W = S.train_som(X[:n], W0, xd, K // xd, rlen=1)
    return W.float().contiguous()


def run(kind, C, K, nfov):
    X = pixie_rows(nfov, C)
    n = X.shape[0]
    datasets = [("P", X, trained_codebook(X, K))]
    if kind == "timingU":
        kind = "timing"
        X.uniform_()
        datasets = [("U", X, trained_codebook(X, K))]
    if kind == "parity":
        U = torch.rand((1 << 21, C), device="cuda")
        datasets.append(("U", U, trained_codebook(U, K)))
        datasets.append(("U-raw", U, U[:K].contiguous()))
    for name, D, W in datasets:
        m = D.shape[0]
        stats = torch.zeros(S.NSTATS, dtype=torch.int64, device="cuda")
        lab = torch.empty(m, dtype=torch.int32, device="cuda")
        S.bmu(D, W, labels=lab, stats=stats)
        torch.cuda.synchronize()
        st = stats.cpu().numpy()
        line = (f"[{tag}] {kind} {name} n={m} C={C} K={K}: flagged={st[0]/m:.5f} pairs/row={st[1]/m:.4f} "
                f"fp64={st[2]} fixup={st[3]}")
        if kind in ("parity", "margin"):
            sub = min(m, 1 << 22)
            ref = S.bmu(D[:sub], W, flags=S.FLAG_FORCE_EXACT)
            torch.cuda.synchronize()
            line += f" mismatches_vs_exact({sub})={int((lab[:sub] != ref).sum())}"
        if kind == "timing":
            for _ in range(3):
                S.bmu(D, W, labels=lab)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            e0.record()
            for _ in range(reps):
                S.bmu(D, W, labels=lab)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            gbs = m * (4 * C + 4) / ms / 1e6
            line += f" {ms:.3f} ms {m/ms/1e6:.2f} Gpx/s {gbs:.0f} GB/s frac={gbs/bench.measured_peaks()[0]:.3f}"
        print(line, flush=True)


if __name__ == "__main__":
    kind = sys.argv[1]
    C = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    K = int(sys.argv[3]) if len(sys.argv) > 3 else 100
    nfov = int(sys.argv[4]) if len(sys.argv) > 4 else 4
    run(kind, C, K, nfov)
