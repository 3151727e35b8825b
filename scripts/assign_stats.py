"""Assign-kernel statistics and time for one shape on the bench's data (Pixie-like rows, codebook
trained on their 10 % subset): rows flagged / pairs / fp64 rows / fix-up rows, time, roofline
fraction.  usage: assign_stats.py nfov hw C xdim ydim [delta_scale]"""
import os
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from ark_analysis_b200 import _native, som as S  # noqa: E402

nfov, hw, C, xd, yd = [int(v) for v in sys.argv[1:6]]
dev = torch.device("cuda", 0)
wl = bench.Workload(torch, S, dev, 0, 1, None, nfov, hw, C, xd, yd)
wl.step()
torch.cuda.synchronize()
stats = torch.zeros(S.NSTATS, dtype=torch.int64, device=dev)
S.bmu(wl.X, wl.W32, labels=wl.labels, stats=stats)
torch.cuda.synchronize()
st = stats.cpu().numpy()
n = wl.n
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(2):
    S.bmu(wl.X, wl.W32, labels=wl.labels)
e0.record()
for _ in range(5):
    S.bmu(wl.X, wl.W32, labels=wl.labels)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
peak = bench.measured_peaks()[0]
print(f"C={C} K={xd * yd} n={n}: assign {ms:.3f} ms = {n / ms / 1e6:.2f} Gpx/s, "
      f"{n * (4 * C + 4) / ms / 1e6 / peak:.3f} of roofline; flagged {st[0] / n:.4f} pairs/flagged "
      f"{st[1] / max(st[0], 1):.2f} fp64 rows {st[2] / n:.5f} fix-up rows {st[3] / n:.6f}",
      flush=True)
