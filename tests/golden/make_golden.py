"""Generates tests/golden/*.npz -- small input/output vectors of the ORACLE (oracle/pixie_oracle.c).

The reference holds no golden vectors for this path and its arithmetic dependency (pyFlowSOM) is
not installable here, so these vectors pin the oracle against itself across
rebuilds and give the GPU tests fixed inputs; they are NOT outputs of the reference ("parity
unpinned").  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from conftest import pixie_like  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def bmu_case(name, X, W):
    labels, dists = oracle.map_data_to_nodes_f32(W, X)
    np.savez_compressed(os.path.join(HERE, name), X=X, W=W, labels=labels, dists=dists)
    print(name, X.shape, W.shape, "labels", labels.min(), labels.max())


def main():
    r = np.random.default_rng(2024)
    # cfg1-shaped slice: 16 channels, 10x10 SOM
    X = r.random((4096, 16), dtype=np.float32)
    bmu_case("bmu_u16_k100.npz", X, X[r.choice(4096, 100, replace=False)].copy())
    # Pixie-like, 32 channels
    P = pixie_like(4096, 32)
    bmu_case("bmu_p32_k100.npz", P, P[r.choice(4096, 100, replace=False)].copy())
    # exact ties: duplicated codebook rows and rows equal to nodes (first minimum must win)
    Xt = r.integers(0, 3, (1024, 8)).astype(np.float32)
    Wt = Xt[:40].copy()
    Wt[20:] = Wt[:20]
    bmu_case("bmu_ties_c8_k40.npz", Xt, Wt)
    # ragged channel count, K not a multiple of anything, NaN / Inf rows
    Xr = r.random((777, 15), dtype=np.float32)
    Xr[5, 3] = np.nan
    Xr[9, :] = np.inf
    Xr[11, 0] = -np.inf
    bmu_case("bmu_nan_c15_k49.npz", Xr, Xr[100:149].copy())
    # batch SOM (the algorithm the GPU runs): 2 passes on 3000 rows, 16 channels, 6x5 map
    Xs = pixie_like(3000, 16, seed=7)
    idx = oracle.init_codebook_indices(3000, 30, 42)
    Wb = oracle.som_batch(Xs, 6, 5, rlen=2, alpha_range=(0.05, 0.01), seed=42, init_idx=idx)
    np.savez_compressed(os.path.join(HERE, "som_batch_p16_6x5.npz"), X=Xs, init_idx=idx, W=Wb)
    print("som_batch", Wb.shape, float(Wb.sum()))
    # online SOM restatement (Appendix A, UNVERIFIED): pins determinism of the restatement only
    Wo = oracle.som_online(Xs.astype(np.float64), 6, 5, rlen=1, seed=42, init_idx=idx)
    np.savez_compressed(os.path.join(HERE, "som_online_p16_6x5.npz"), W=Wo)


if __name__ == "__main__":
    main()
