"""Event timeline of CTA 0 for two consecutive training steps (diagnostic build: make prof,
PIXIE_LIB_PATH=.../libpixie_b200_prof.so, PIXIE_TRACE_STEP=<first step>).
usage: trace_train_step.py n C xdim ydim"""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from ark_analysis_b200 import _native, som as S  # noqa: E402
from conftest import pixie_like  # noqa: E402

n, C, xd, yd = [int(v) for v in sys.argv[1:5]]
K = xd * yd
base = torch.from_numpy(pixie_like(1 << 20, C)).cuda()
X = base.repeat((n + base.shape[0] - 1) // base.shape[0], 1)[:n].contiguous()
W0 = X[:K].to(torch.float64)
L = _native.lib()
buf = np.zeros((2 << 15) + 512, np.uint64)
for _ in range(3):
    S.train_som(X, W0, xd, yd, rlen=1, batches_per_pass=32)
    torch.cuda.synchronize()
    cnt = L.pixie_debug_trace(buf.ctypes.data_as(ctypes.c_void_p), (1 << 15) + 256)
ev = buf[:2 * cnt].reshape(-1, 2)
ev = ev[ev[:, 1] != 0]
cnt = len(ev)
key, t = ev[:, 0], ev[:, 1].astype(np.int64)
st, warp, e, seq = (key >> 48).astype(int), ((key >> 40) & 0xff).astype(int), \
    ((key >> 32) & 0xff).astype(int), (key & 0xffffffff).astype(int)
arr = buf[2 << 15:].astype(np.int64).reshape(2, 256)
for k in range(2):
    a = arr[k][arr[k] != 0]
    if len(a):
        a = (a - a.min()) / 1e3
        srt = np.sort(a)
        print(f"step +{k}: CTA arrival at the end of the tiles, us after the first: median {np.median(a):.2f} "
              f"p90 {srt[int(0.9 * len(a))]:.2f} max {a.max():.2f}; CTAs 0-95 (9 tiles) mean {a[:96].mean():.2f}, "
              f"96-147 (8 tiles) mean {a[96:].mean():.2f}; slowest CTAs {np.argsort(a)[-6:].tolist()}")
t0 = t.min()
names = {0: "tma issued", 1: "mma: X seen", 2: "mma: committed", 3: "epi: X seen", 4: "epi: acc ready",
         5: "epi: passes done", 6: "epi: labels done", 7: "epi: group met", 8: "epi: tile done",
         9: "step_finish in", 10: "step_finish out"}
order = np.argsort(t, kind="stable")
print(f"{cnt} events; step, warp, event, tile seq, us since first event")
for i in order:
    print(f"st {st[i]:2d} w{warp[i]:2d} {names.get(e[i], e[i]):18s} seq {seq[i]:5d} {(t[i] - t0) / 1e3:9.2f}")
