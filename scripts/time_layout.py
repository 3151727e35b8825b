"""Timing of the N2 kernel (column buffers -> device matrix) against its HBM roofline:
12 C bytes per pixel (8 C read as float64 columns, 4 C written as fp32 rows)."""
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from ark_analysis_b200 import som as S  # noqa: E402

peak = bench.measured_peaks()[0]
for C, n in ((32, 16 << 20), (16, 32 << 20), (40, 12 << 20), (100, 5 << 20)):
    cols = torch.rand((C, n), dtype=torch.float64, device="cuda")
    div = torch.rand(C, dtype=torch.float64, device="cuda") + 0.5
    out = torch.zeros((n, (C + 3) // 4 * 4), dtype=torch.float32, device="cuda")
    for with_div in (True, False):
        for _ in range(3):
            S.columns_to_rows(cols, div if with_div else None, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            S.columns_to_rows(cols, div if with_div else None, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        gbs = n * 12 * C / ms / 1e6
        print(f"columns_to_rows C={C} n={n} divide={with_div}: {ms:.3f} ms {n/ms/1e6:.2f} Gpx/s "
              f"{gbs:.0f} GB/s frac={gbs/peak:.3f}", flush=True)
    del cols, out
