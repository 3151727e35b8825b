import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from ark_analysis_b200 import som as S
from conftest import pixie_like
n, C, K = 5241600, 32, 100
base = torch.from_numpy(pixie_like(1 << 20, C)).cuda()
X = base.repeat(5, 1)[:n].contiguous()
W0 = X[:K].to(torch.float64)
rlen = int(sys.argv[2]) if len(sys.argv) > 2 else 1
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    W = S.train_som(X, W0, 10, 10, rlen=rlen, batches_per_pass=32)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    W = S.train_som(X, W0, 10, 10, rlen=rlen, batches_per_pass=32)
e1.record(); torch.cuda.synchronize()
print("train ms per pass", e0.elapsed_time(e1) / 10 / rlen, "rlen", rlen)

# phase timers of CTA 0 (CodebookAux.phase_ns): workspace offset of the control block + 64 bytes
ws = list(S._ws_cache.values())[0]
off = 5 * 512 * 128
ph = ws[off + 56: off + 56 + 48].view(torch.int64).cpu().numpy()
names = ["tiles", "barrier1", "fold", "barrier2", "update", "barrier3"]
print("per-step us (last launch):", {n: round(v / 32 / rlen / 1e3, 2) for n, v in zip(names, ph)})
