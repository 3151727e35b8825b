"""Assign path on the bench workload (Pixie-like rows, trained codebook) for ncu.
Usage: prof_assign_p.py [nfov] [C] [K] [reps]"""
import sys

import torch

sys.path.insert(0, ".")
from scripts.variant_experiment import pixie_rows, trained_codebook  # noqa: E402
from ark_analysis_b200 import som as S  # noqa: E402

nfov = int(sys.argv[1]) if len(sys.argv) > 1 else 50
C = int(sys.argv[2]) if len(sys.argv) > 2 else 32
K = int(sys.argv[3]) if len(sys.argv) > 3 else 100
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
X = pixie_rows(nfov, C)
W = trained_codebook(X, K)
lab = torch.empty(X.shape[0], dtype=torch.int32, device="cuda")
for _ in range(reps):
    S.bmu(X, W, labels=lab)
torch.cuda.synchronize()
print("done", X.shape, K)
