"""One training pass on the step-by-step path for ncu's launch list: argv = C xdim n_train."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from ark_analysis_b200 import som as S  # noqa: E402
from conftest import pixie_like  # noqa: E402

C, xd, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
K = xd * xd
base = torch.from_numpy(pixie_like(1 << 20, C)).cuda()
X = base.repeat((n + (1 << 20) - 1) >> 20, 1)[:n].contiguous()
idx = np.random.default_rng(42).choice(1 << 20, K, replace=False)
W0 = base[torch.from_numpy(idx).cuda()].double()
for _ in range(2):
    W = S.train_som(X, W0, xd, xd, rlen=1, batches_per_pass=32)
torch.cuda.synchronize()
print("done")
