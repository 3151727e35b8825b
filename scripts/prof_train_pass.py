"""Training-pass time and the per-step phase timers of CTA 0 (CodebookAux.phase_ns) for one shape.
usage: prof_train_pass.py [n C xdim ydim [reps [rlen]]]   (default: the cfg2 training subset)"""
import sys

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from ark_analysis_b200 import som as S  # noqa: E402
from conftest import pixie_like  # noqa: E402

a = sys.argv[1:]
n, C, xd, yd = (int(a[0]), int(a[1]), int(a[2]), int(a[3])) if len(a) >= 4 else (5241600, 32, 10, 10)
reps = int(a[4]) if len(a) > 4 else 10
rlen = int(a[5]) if len(a) > 5 else 1
K = xd * yd
base = torch.from_numpy(pixie_like(1 << 20, C)).cuda()
X = base.repeat((n + base.shape[0] - 1) // base.shape[0], 1)[:n].contiguous()
W0 = X[:K].to(torch.float64)
for _ in range(3):
    W = S.train_som(X, W0, xd, yd, rlen=rlen, batches_per_pass=32)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    W = S.train_som(X, W0, xd, yd, rlen=rlen, batches_per_pass=32)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps / rlen
print(f"n={n} C={C} K={K}: train {ms:.3f} ms per pass = {n / ms / 1e6:.2f} Gpx/s, "
      f"{ms * 1e3 / 32:.1f} us per step", flush=True)

# phase timers: workspace offset of the control block (behind the codebook image) + 56 bytes
ws = list(S._ws_cache.values())[0]
off = 5 * 512 * 128
ph = ws[off + 56: off + 56 + 48].view(torch.int64).cpu().numpy()
names = ["tiles", "combine+barrier1", "fold(+exchange)", "barrier2", "update", "barrier3"]
print("  per-step us (last launch):", {k: round(float(v) / 32 / rlen / 1e3, 2) for k, v in zip(names, ph)})
tc = ws[off + 104: off + 104 + 64].view(torch.int64).cpu().numpy()
if tc[6] > 0:
    tn = ["wait X", "norm", "acc wait + passes", "resolve", "fix-ups", "accumulate work"]
    print(f"  tile phases of warp 0 / CTA 0, us per tile over {tc[6]} tiles:",
          {k: round(float(v) / float(tc[6]) / 1965.0, 2) for k, v in zip(tn, tc[:6])},
          "group wait before accumulate", round(float(tc[7]) / float(tc[6]) / 1965.0, 2))
