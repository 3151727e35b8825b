"""How much room the candidate window of the split-operand assign kernel has: labels against the
exact fp64 kernel while PIXIE_DELTA_SCALE shrinks the window, on the bench data and on three
harder data sets (large values, mixed signs, pairs of near-identical nodes).  The first scale at
which labels differ is where the real score error sits relative to the bound.
usage: x3_margin.py [nfov]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from ark_analysis_b200 import som as S  # noqa: E402

nfov = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
wl = bench.Workload(torch, S, dev, 0, 1, None, nfov, 1024, 32, 10, 10)
wl.step()
torch.cuda.synchronize()


def sweep(name, X, W):
    ref = S.bmu(X, W, flags=S.FLAG_FORCE_EXACT)
    n = X.shape[0]
    for k in range(0, 15, 2):
        os.environ["PIXIE_DELTA_SCALE"] = repr(2.0 ** -k)
        stats = torch.zeros(S.NSTATS, dtype=torch.int64, device=dev)
        lab = S.bmu(X, W, flags=S.FLAG_FORCE_TC, stats=stats)
        torch.cuda.synchronize()
        bad = int((lab != ref).sum())
        print(f"{name:28s} window x 2^-{k:<2d}: rows flagged {int(stats[0]) / n:9.6f}  "
              f"labels differing from the exact kernel {bad}", flush=True)
    os.environ.pop("PIXIE_DELTA_SCALE")


sweep("bench data (cfg2 rows)", wl.X, wl.W32)
g = torch.Generator(device=dev).manual_seed(5)
X = wl.X[: 1 << 22]
sweep("values x 1000", (X * 1000.0).contiguous(), (wl.W32 * 1000.0).contiguous())
Xs = (X - 0.05).contiguous()
sweep("mixed signs", Xs, (wl.W32 - 0.05).contiguous())
W2 = wl.W32.clone()
W2[1::2] = W2[::2] * (1 + 2.0 ** -18)
sweep("near-identical node pairs", X.contiguous(), W2.contiguous())
U = torch.rand((1 << 22, 32), generator=g, device=dev)
sweep("uniform rows, nodes = rows", S.to_device_matrix(U), U[:100].contiguous())
