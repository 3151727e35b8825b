"""N4 timing: label histogram and cluster-mask scatter on cfg2-sized label arrays (50 FOVs x
1024 x 1024 pixels, blob-like cells, K = 100), CUDA events, GB/s against the measured HBM peak."""
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from ark_analysis_b200 import som as S  # noqa: E402

peak = bench.measured_peaks()[0]
H = W = 1024
nf = 50
n = nf * H * W
g = torch.Generator(device="cuda").manual_seed(1)
# cells as 16 x 16 blocks (image order: runs of 16 equal labels), ~4000 cells per FOV
yy = torch.arange(H, device="cuda").view(H, 1) // 16
xx = torch.arange(W, device="cuda").view(1, W) // 16
seg1 = (yy * (W // 16) + xx + 1).to(torch.int32).reshape(-1)
seg = seg1.repeat(nf)
clu = torch.randint(1, 101, (n,), device="cuda", generator=g, dtype=torch.int32)
# spatially coherent clusters (what a SOM produces on blurred images): 4-pixel runs
clu_runs = torch.randint(1, 101, (n // 4,), device="cuda", generator=g,
                         dtype=torch.int32).repeat_interleave(4)
n_seg = int(seg.max()) + 1


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


counts = torch.zeros((n_seg, 101), dtype=torch.int32, device="cuda")
for name, c in (("iid clusters", clu), ("4-pixel runs", clu_runs)):
    ms = timed(lambda: S.label_histogram(seg, c, n_seg, 101, counts=counts))
    gbs = n * 8 / ms / 1e6
    print(f"label_histogram {name}: n={n} {ms:.3f} ms {n/ms/1e6:.1f} Gpx/s {gbs:.0f} GB/s "
          f"frac={gbs/peak:.3f} (8 B/pixel)")
rows = (torch.arange(H, device="cuda", dtype=torch.int32).view(H, 1).expand(H, W)).reshape(-1)
cols = (torch.arange(W, device="cuda", dtype=torch.int32).view(1, W).expand(H, W)).reshape(-1)
lut = torch.arange(101, dtype=torch.int16)
k1 = clu[:H * W].contiguous()
for uniq in (True, False):
    ms = timed(lambda: S.scatter_labels(rows, cols, k1, H, W, id_map=lut, unique=uniq))
    b = H * W * (14 if uniq else 14 + 12)
    print(f"scatter_labels unique={uniq}: n={H*W} {ms:.3f} ms {H*W/ms/1e6:.2f} Gpx/s "
          f"{b/ms/1e6:.0f} GB/s frac={b/ms/1e6/peak:.3f}")
