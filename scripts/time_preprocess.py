"""N3 timing: preprocessing of one cfg2-sized FOV (1024 x 1024 x 32 fp32 image -> fp32 SOM rows) on
the device, CUDA events.  Algorithmic bytes per pixel-channel: 4 in + 4 out (fp32 rows only) or
4 + 12 (fp32 + fp64 rows)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from ark_analysis_b200 import pixie_preprocessing as PP  # noqa: E402

peak = bench.measured_peaks()[0]
H = W = 1024
for C in (16, 32, 40):
    img = torch.empty((H, W, C), device="cuda").exponential_(1.0)
    norm = np.linspace(0.5, 2.0, C)
    for x64 in (False, True):
        def run():
            return PP.preprocess_fov_device(img, norm, 0.5 * C, 2, want_x64=x64)
        for _ in range(2):
            out = run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            out = run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        b = H * W * C * 4 + out["n"] * C * (12 if x64 else 4)
        print(f"preprocess_fov C={C} x64={x64}: kept {out['n']} of {H*W}; {ms:.3f} ms "
              f"{H*W/ms/1e6:.2f} Gpx/s; algorithmic {b/ms/1e6:.0f} GB/s frac={b/ms/1e6/peak:.3f}")
