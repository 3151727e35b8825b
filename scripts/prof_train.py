"""Runs the training path a few times on the cfg2 training subset (for an ncu launch list)."""
import sys
import torch
sys.path.insert(0, ".")
from ark_analysis_b200 import som as S
n, C = 5241600, 32
X = torch.rand((n, C), device="cuda")
W0 = X[:100].to(torch.float64)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    W = S.train_som(X, W0, 10, 10, rlen=1, batches_per_pass=32)
torch.cuda.synchronize()
print("done")
