"""BASELINE.json configs[4]: BMU throughput sweep N x C x K at 1/2/4/8 GPUs with the CPU path beside it.

N pixels are sharded over the ranks (rows are independent: no collective on the data path), every
rank assigns its N / world rows, the time is the max over ranks (device events, barrier on both
sides), the value the whole job's pixels/s.  Rows are "U" (uniform [0,1)) or "P" (Pixie-like, the
bench generator: prototype + noise, row-normalised, channel-normalised).  The codebook is trained
for one pass on the first 2^20 rows of rank 0 and broadcast.  The CPU column is the C restatement of
pyFlowSOM's map_data_to_nodes (oracle/pixie_oracle.c, the reference's call at cluster_helpers.py:152-157)
on all host threads over a bounded sample of the same rows, timed by rank 0 only.

usage: [torchrun ...] python scripts/sweep_cfg5.py out.md   (env: SWEEP_C, SWEEP_K, SWEEP_N, SWEEP_DIST)
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from ark_analysis_b200 import som as S  # noqa: E402

out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/sweep_cfg5.md"
rank = int(os.environ.get("RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
Cs = [int(v) for v in os.environ.get("SWEEP_C", "16,32,64").split(",")]
Ks = [int(v) for v in os.environ.get("SWEEP_K", "100,400").split(",")]
Ns = [int(float(v)) for v in os.environ.get("SWEEP_N", "1e6,1e7,1e8,1e9").split(",")]
Ds = os.environ.get("SWEEP_DIST", "U,P").split(",")
peak = bench.measured_peaks()[0]
cores = os.cpu_count() or 1
CPU_ROWS = 1 << 20

rows = [f"| N | C | K | rows | GPUs | ms | Gpx/s (all GPUs) | GB/s per GPU | frac of {peak:.0f} GB/s | "
        f"rows rechecked | CPU Mpx/s ({cores} threads) | note |", "|" + "---|" * 12]


def fill(X, kind, seed):
    n, C = X.shape
    if kind == "U":
        g = torch.Generator(device=dev).manual_seed(seed)
        for i in range(0, n, 1 << 26):
            X[i:i + (1 << 26)].uniform_(generator=g)
        return
    hw = 1024
    npx = hw * hw
    for i in range(0, n, 8 * npx):
        m = min(n - i, 8 * npx)
        nf = (m + npx - 1) // npx
        buf = torch.empty((nf * npx, C), device=dev, dtype=torch.float32)
        bench.gen_fovs_device(torch, dev, [seed * 4096 + i // npx + f for f in range(nf)], buf, hw, C)
        X[i:i + m] = buf[:m]
        del buf


cpu_cache = {}
for C in Cs:
    for K in Ks:
        xd = int(round(np.sqrt(K)))
        for kind in Ds:
            for N in Ns:
                note = ""
                n = (N + world - 1) // world
                free = torch.cuda.mem_get_info()[0]
                cap = int(0.8 * free / (4 * C + 4)) // 128 * 128
                if n > cap:
                    note = f"per-GPU rows capped to {cap} (HBM)"
                    n = cap
                X = torch.empty((n, C), device=dev, dtype=torch.float32)
                fill(X, kind, 1000 * rank + 1)
                m = min(n, 1 << 20) // 128 * 128
                W = torch.empty((K, C), device=dev, dtype=torch.float32)
                if rank == 0:
                    idx = np.random.default_rng(42).choice(m, K, replace=False)
                    W0 = X[torch.from_numpy(idx).to(dev)].double()
                    W.copy_(S.train_som(X[:m], W0, xd, K // xd, rlen=1).float())
                if dist is not None:
                    dist.broadcast(W, 0)
                lab = torch.empty(n, dtype=torch.int32, device=dev)
                stats = torch.zeros(S.NSTATS, dtype=torch.int64, device=dev)
                S.bmu(X, W, labels=lab, stats=stats)
                torch.cuda.synchronize()
                st = stats.cpu().numpy()
                reps = 5 if N <= 10**8 else 2
                for _ in range(2):
                    S.bmu(X, W, labels=lab)
                torch.cuda.synchronize()
                if dist is not None:
                    dist.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    S.bmu(X, W, labels=lab)
                e1.record()
                torch.cuda.synchronize()
                ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
                tot = torch.tensor([float(n)], device=dev, dtype=torch.float64)
                if dist is not None:
                    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
                    dist.all_reduce(tot)
                ms = float(ms)
                total_rows = float(tot)
                if rank == 0:
                    key = (C, K, kind)
                    if key not in cpu_cache:
                        import oracle
                        xs = X[:min(n, CPU_ROWS)].cpu().numpy().astype(np.float64)
                        w64 = W.cpu().numpy().astype(np.float64)
                        oracle.map_data_to_nodes_mt(w64, xs[:65536], cores)  # warm the thread pool
                        t0 = time.perf_counter()
                        oracle.map_data_to_nodes_mt(w64, xs, cores)
                        cpu_cache[key] = xs.shape[0] / (time.perf_counter() - t0) / 1e6
                    gbs = n * (4 * C + 4) / ms / 1e6
                    rows.append(f"| {N:.0e} | {C} | {K} | {kind} | {world} | {ms:.3f} | "
                                f"{total_rows / ms / 1e6:.2f} | {gbs:.0f} | {gbs / peak:.3f} | "
                                f"{st[0] / n:.4f} | {cpu_cache[key]:.2f} | {note} |")
                    print(rows[-1], flush=True)
                del X, lab
                torch.cuda.empty_cache()
if rank == 0:
    os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
    with open(out, "w") as f:
        f.write("\n".join(rows) + "\n")
if dist is not None:
    dist.destroy_process_group()
