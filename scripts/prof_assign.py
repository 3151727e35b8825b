"""Runs the assign path a few times on one config (for ncu).
Usage: prof_assign.py nfov hw C K reps [kind]   kind: U (uniform) or S (sparse, few candidates)"""
import sys
import torch
sys.path.insert(0, ".")
from ark_analysis_b200 import som as S
nfov, hw, C, K, reps = [int(a) for a in sys.argv[1:6]]
n = nfov * hw * hw
X = torch.rand((n, C), device="cuda")
W = X[torch.randperm(n, device="cuda")[:K]].contiguous()
lab = torch.empty(n, dtype=torch.int32, device="cuda")
for _ in range(reps):
    S.bmu(X, W, labels=lab)
torch.cuda.synchronize()
print("done", n, C, K)
