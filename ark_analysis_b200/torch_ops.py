"""``torch.library`` registration of the device operators: the same
C-ABI calls ``ark_analysis_b200.som`` makes, reachable as ``torch.ops.pixie_b200.*`` -- for callers
that want the kernels inside a torch graph (``torch.compile`` sees them as opaque ops with the
shapes declared below; nothing here is compiled by torch).

    labels = torch.ops.pixie_b200.bmu(X, W)                     # int32 [n], 1-indexed
    labels, SN = torch.ops.pixie_b200.bmu_sums(X, W)            # + per-node sums / counts [K, C+1]
    W64 = torch.ops.pixie_b200.som_train(X, W0, xdim, ydim, rlen, lr_start, lr_end, batches)
"""
import torch

from . import som

__all__ = ["bmu", "bmu_sums", "som_train"]


@torch.library.custom_op("pixie_b200::bmu", mutates_args=(), device_types="cuda")
def bmu(X: torch.Tensor, W: torch.Tensor) -> torch.Tensor:
    return som.bmu(X, W)


@bmu.register_fake
def _(X, W):
    return X.new_empty((X.shape[0],), dtype=torch.int32)


@torch.library.custom_op("pixie_b200::bmu_sums", mutates_args=(), device_types="cuda")
def bmu_sums(X: torch.Tensor, W: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    labels, SN = som.bmu(X, W, want_sums=True)
    return labels, SN


@bmu_sums.register_fake
def _(X, W):
    return (X.new_empty((X.shape[0],), dtype=torch.int32),
            X.new_empty((W.shape[0], X.shape[1] + 1), dtype=torch.float64))


@torch.library.custom_op("pixie_b200::som_train", mutates_args=(), device_types="cuda")
def som_train(X: torch.Tensor, W0: torch.Tensor, xdim: int, ydim: int, rlen: int, lr_start: float,
              lr_end: float, batches_per_pass: int) -> torch.Tensor:
    return som.train_som(X, W0, xdim, ydim, rlen=rlen, alpha_range=(lr_start, lr_end),
                         batches_per_pass=batches_per_pass if batches_per_pass > 0 else None)


@som_train.register_fake
def _(X, W0, xdim, ydim, rlen, lr_start, lr_end, batches_per_pass):
    return X.new_empty((xdim * ydim, X.shape[1]), dtype=torch.float64)
