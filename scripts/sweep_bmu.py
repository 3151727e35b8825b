"""BASELINE.json configs[4]: BMU throughput sweep N x C x K on one GPU (uniform rows U[0,1), codebook
= K rows of the data trained for one pass on the first 2^20 rows).  Writes a markdown table.
Usage: python scripts/sweep_bmu.py [out.md] [max_N]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from ark_analysis_b200 import som as S  # noqa: E402

out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/sweep.md"
max_n = int(float(sys.argv[2])) if len(sys.argv) > 2 else int(1e9)
peak = bench.measured_peaks()[0]
free = torch.cuda.mem_get_info()[0]
rows = ["| N | C | K | kernel | ms | Gpx/s | GB/s | frac of %.0f GB/s | rows rechecked | note |" % peak,
        "|---|---|---|---|---|---|---|---|---|---|"]
import os  # noqa: E402
only_c = [int(v) for v in os.environ.get("SWEEP_C", "16,32,64").split(",")]
only_k = [int(v) for v in os.environ.get("SWEEP_K", "100,400").split(",")]
for C in only_c:
    for K in only_k:
        for N in (10**6, 10**7, 10**8, 10**9):
            if N > max_n:
                continue
            note = ""
            n = N
            cap = int(0.8 * free / (4 * C + 4)) // 128 * 128
            if n > cap:
                note = f"resident-capped to {cap} rows"
                n = cap
            xd = int(round(np.sqrt(K)))
            X = torch.empty((n, C), device="cuda", dtype=torch.float32)
            g = torch.Generator(device="cuda").manual_seed(42)
            step = 1 << 26
            for i in range(0, n, step):
                X[i:i + step].uniform_(generator=g)
            m = min(n, 1 << 20) // 128 * 128
            idx = np.random.default_rng(42).choice(m, K, replace=False)
            W0 = X[torch.from_numpy(idx).cuda()].double()
            W = S.train_som(X[:m], W0, xd, K // xd, rlen=1).float().contiguous()
            lab = torch.empty(n, dtype=torch.int32, device="cuda")
            stats = torch.zeros(S.NSTATS, dtype=torch.int64, device="cuda")
            S.bmu(X, W, labels=lab, stats=stats)
            torch.cuda.synchronize()
            st = stats.cpu().numpy()
            kern = {1: "tensor-core", 2: "exact fp64"}.get(int(st[4]), "?")
            if kern != "tensor-core" and n > 10**6:
                rows.append(f"| {N:.0e} | {C} | {K} | {kern} | - | - | - | - | - | no tensor-core plan "
                            f"for this shape (codebook image > shared memory); skipped above 1e6 |")
                del X, lab
                continue
            reps = 5 if n <= 10**8 else 2
            for _ in range(2):
                S.bmu(X, W, labels=lab)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                S.bmu(X, W, labels=lab)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            gbs = n * (4 * C + 4) / ms / 1e6
            rows.append(f"| {N:.0e} | {C} | {K} | {kern} | {ms:.3f} | {n/ms/1e6:.2f} | {gbs:.0f} | "
                        f"{gbs/peak:.3f} | {st[0]/n:.3f} | {note} |")
            print(rows[-1], flush=True)
            del X, lab
            torch.cuda.empty_cache()
open(out, "w").write("\n".join(rows) + "\n")
