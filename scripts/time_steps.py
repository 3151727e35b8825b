"""Per-step time and recheck statistics of a training pass driven through the step API."""
import sys, torch, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from ark_analysis_b200 import som as S
from conftest import pixie_like
n, C, K, B = 5241600, 32, 100, 32
base = torch.from_numpy(pixie_like(1 << 20, C)).cuda()
X = base.repeat(5, 1)[:n].contiguous()
W64 = X[:K].to(torch.float64).clone()
W32 = torch.empty((K, C), dtype=torch.float32, device="cuda")
SN = torch.zeros((K, C + 1), dtype=torch.float64, device="cuda")
S.som_apply(W64, W32, SN, 10, 10, 1.0, 0.0)
rr = S.default_radius(10, 10)
for t in range(B):
    stats = torch.zeros(S.NSTATS, dtype=torch.int64, device="cuda")
    S.som_accum(X, W32, t % B, B, SN=SN, stats=stats)  # warm
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stats.zero_()
    e0.record(); S.som_accum(X, W32, t % B, B, SN=SN, stats=stats); e1.record()
    torch.cuda.synchronize()
    st = stats.cpu().numpy(); rows = n // B
    print(f"step {t:2d}: {e0.elapsed_time(e1)*1e3:7.1f} us flagged {st[0]/rows:.3f} pairs/row {st[1]/rows:.2f} fp64 {st[2]} fixup {st[3]}")
    sigma, alpha = S.step_schedule(t, B, (0.05, 0.01), rr)
    S.som_apply(W64, W32, SN, 10, 10, sigma, alpha)
