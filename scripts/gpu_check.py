"""Exploratory GPU check (not a test): parity of each kernel against the oracle, recheck
statistics and first timings.  Usage: python scripts/gpu_check.py <stage> [...]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import oracle  # noqa: E402
from ark_analysis_b200 import som as S  # noqa: E402


def pixie_like(n, C, seed=12345, nproto=30):
    rng = np.random.default_rng(seed)
    protos = rng.dirichlet(np.full(C, 0.3), size=nproto)
    which = rng.integers(0, nproto, n)
    X = protos[which] + np.abs(rng.normal(0, 0.05, (n, C)))
    X = np.maximum(X, 0)
    X /= X.sum(1, keepdims=True)
    X /= np.quantile(X, 0.999, axis=0)
    return X.astype(np.float32)


def data(kind, n, C, seed=0):
    if kind == "U":
        return np.random.default_rng(seed).random((n, C), dtype=np.float32)
    return pixie_like(n, C, seed + 12345)


def check_bmu(kind, n, C, K, flags, trained=False):
    X = data(kind, n, C)
    rng = np.random.default_rng(1)
    W = X[rng.choice(n, K, replace=False)].copy()
    if trained:
        xd = int(round(np.sqrt(K)))
        W = oracle.som_batch(X[:20000], xd, K // xd, rlen=1).astype(np.float32)
    ref, _ = oracle.map_data_to_nodes_f32(W, X)
    Xd = S.to_device_matrix(X)
    Wd = torch.from_numpy(W).cuda()
    stats = torch.zeros(S.NSTATS, dtype=torch.int64, device="cuda")
    lab = S.bmu(Xd, Wd, flags=flags, stats=stats)
    torch.cuda.synchronize()
    lab = lab.cpu().numpy()
    bad = int((lab != ref).sum())
    st = stats.cpu().numpy()
    print(f"bmu {kind} n={n} C={C} K={K} flags={flags} trained={trained}: mismatches={bad} "
          f"flagged={st[0]/n:.4f} pairs/row={st[1]/n:.3f} fp64rows={st[2]} fixup={st[3]} kernel={st[4]}",
          flush=True)
    if bad:
        idx = np.nonzero(lab != ref)[0][:10]
        print("   first bad rows", idx, "got", lab[idx], "want", ref[idx], flush=True)
    return bad


def stage_exact():
    for C, K in [(16, 100), (7, 30)]:
        check_bmu("U", 5000, C, K, S.FLAG_FORCE_EXACT)


def stage_tc_small():
    check_bmu("U", 128 * 3, 32, 100, S.FLAG_FORCE_TC)
    check_bmu("U", 20000, 32, 100, S.FLAG_FORCE_TC)
    check_bmu("P", 20000, 32, 100, S.FLAG_FORCE_TC)


def stage_tc_shapes():
    for kind in ("U", "P"):
        for C, K in [(16, 100), (32, 100), (40, 400), (100, 100), (15, 200), (64, 100), (8, 16),
                     (128, 256), (33, 49)]:
            try:
                check_bmu(kind, 30000 + 77, C, K, S.FLAG_FORCE_TC, trained=(K in (100, 400)))
            except Exception as e:  # noqa: BLE001
                print("   FAILED", kind, C, K, repr(e), flush=True)


def stage_timing():
    for (nfov, hw, C, K) in [(50, 1024, 32, 100), (12, 2048, 40, 400), (1, 2236, 100, 100)]:
        n = nfov * hw * hw
        Xd = torch.rand((n, C), device="cuda", dtype=torch.float32)
        Wd = Xd[torch.randperm(n, device="cuda")[:K]].contiguous()
        lab = torch.empty(n, dtype=torch.int32, device="cuda")
        stats = torch.zeros(S.NSTATS, dtype=torch.int64, device="cuda")
        for _ in range(2):
            S.bmu(Xd, Wd, labels=lab, stats=stats)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        reps = 5
        for _ in range(reps):
            S.bmu(Xd, Wd, labels=lab)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / reps
        gbs = n * (4 * C + 4) / ms / 1e6
        st = stats.cpu().numpy()
        print(f"timing n={n} C={C} K={K}: {ms:.3f} ms  {n/ms/1e6:.2f} Gpx/s  {gbs:.0f} GB/s "
              f"({gbs/6548.2:.3f} of 6548) flagged={st[0]/n/2:.4f} fixup={st[3]}", flush=True)
        # spot parity on a slice
        sl = slice(0, 200000)
        ref, _ = oracle.map_data_to_nodes_f32(Wd.cpu().numpy(), Xd[sl].cpu().numpy())
        print("   parity on 200k rows: mismatches =", int((lab[sl].cpu().numpy() != ref).sum()),
              flush=True)
        del Xd, lab


def stage_train():
    for kind, n, C, xd, yd in [("U", 20000, 16, 10, 10), ("P", 50000, 32, 10, 10)]:
        X = data(kind, n, C)
        idx = oracle.init_codebook_indices(n, xd * yd, 42)
        t = time.time()
        Wref = oracle.som_batch(X, xd, yd, rlen=1, init_idx=idx)
        tcpu = time.time() - t
        Xd = S.to_device_matrix(X)
        W0 = X[idx].astype(np.float64)
        W = S.train_som(Xd, W0, xd, yd, rlen=1)
        torch.cuda.synchronize()
        t = time.time()
        W = S.train_som(Xd, W0, xd, yd, rlen=1)
        torch.cuda.synchronize()
        tg = time.time() - t
        W = W.cpu().numpy()
        rel = np.abs(W - Wref).max() / np.abs(Wref).max()
        print(f"train {kind} n={n} C={C} {xd}x{yd}: max rel diff {rel:.3e}  cpu {tcpu:.2f}s gpu {tg*1e3:.1f}ms",
              flush=True)


def stage_host():
    X = data("U", 300000, 32)
    W = X[:100].copy()
    ref, dref = oracle.map_data_to_nodes_f32(W, X)
    lab, d = S.map_data_to_nodes(W, X, chunk_rows=65536)
    print("host f32: mismatches", int((lab != ref).sum()), "max dist diff", np.abs(d - dref).max(), flush=True)
    lab, d = S.map_data_to_nodes(W.astype(np.float64), X.astype(np.float64))
    print("host f64: mismatches", int((lab != ref).sum()), "max dist diff", np.abs(d - dref).max(), flush=True)
    Xn = X[:1000].copy()
    Xn[5, 3] = np.nan
    Xn[7, :] = np.inf
    ref, _ = oracle.map_data_to_nodes_f32(W, Xn)
    lab, _ = S.map_data_to_nodes(W, Xn)
    print("NaN rows: mismatches", int((lab != ref).sum()), "labels at 5,7:", lab[5], lab[7], ref[5], ref[7], flush=True)


if __name__ == "__main__":
    for st in sys.argv[1:]:
        print("=== stage", st, flush=True)
        globals()["stage_" + st]()
