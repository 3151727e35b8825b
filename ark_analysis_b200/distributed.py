"""Multi-GPU plumbing of the Pixie SOM path: one process per GPU (``torch.distributed``, NCCL).

The path shards by rows:

* assignment -- FOVs are independent: each rank labels a contiguous range of FOVs against the same
  codebook; no collective on the data path.
* training   -- the batch SOM's per-step statistics (per-node channel sums and counts, K x (C+1)
  float64, <= 131 KB) are summed over ranks with ONE all-reduce per step; every rank then applies
  the identical update.  Mini-batch m of a pass is "all tiles whose GLOBAL index is congruent to m
  mod B", so the trained codebook does not depend on how many ranks the rows are spread over
  (up to float64 summation order).

Everything here is host-side index logic plus the step loop; it is exercised on CPU with the
``gloo`` backend in tests/test_distributed_cpu.py (with the oracle standing in for the kernels)
and on GPUs by ``som.train_som(..., group=...)`` and ``bench.py --gpus N``.
"""
from typing import Callable, List, Sequence, Tuple

TILE = 128


def fov_shards(num_fovs: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous FOV ranges [lo, hi) per rank, sizes differing by at most one FOV."""
    base, extra = divmod(num_fovs, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def tile_aligned_row_shards(n_rows: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous row ranges [lo, hi) per rank whose boundaries are multiples of TILE rows, as
    even as the tile granularity allows.  Training shards must be tile aligned so that a local
    tile is a global tile."""
    ntiles = -(-n_rows // TILE)
    base, extra = divmod(ntiles, world)
    out, t = [], 0
    for r in range(world):
        cnt = base + (1 if r < extra else 0)
        lo, hi = t * TILE, min((t + cnt) * TILE, n_rows)
        out.append((min(lo, n_rows), hi))
        t += cnt
    return out


def first_local_tile(m: int, batches_per_pass: int, tile_offset: int) -> int:
    """First LOCAL tile index of mini-batch m for a shard whose first tile is global tile
    ``tile_offset``: the smallest t >= 0 with (t + tile_offset) % B == m."""
    return (m - tile_offset) % batches_per_pass


def step_schedule(t: int, T: int, alpha_range: Sequence[float], radius_range: Sequence[float]):
    """(sigma, alpha) of step t of T."""
    frac = t / T
    r = radius_range[0] - (radius_range[0] - radius_range[1]) * frac
    r_eff = 0.5 if r < 1.0 else r
    alpha = alpha_range[0] - (alpha_range[0] - alpha_range[1]) * frac
    return 0.5 * r_eff, alpha


def run_training_steps(rlen: int, batches_per_pass: int, tile_offset: int,
                       alpha_range: Sequence[float], radius_range: Sequence[float],
                       accum: Callable[[int, int], object],
                       allreduce: Callable[[object], None],
                       apply: Callable[[object, float, float], None]) -> int:
    """The step loop shared by every backend.

    ``accum(first_tile, stride)`` returns this rank's statistics for the mini-batch,
    ``allreduce(stats)`` sums them over ranks in place, ``apply(stats, sigma, alpha)`` updates the
    codebook.  Returns the number of steps run."""
    B = int(batches_per_pass)
    T = int(rlen) * B
    for t in range(T):
        m = t % B
        stats = accum(first_local_tile(m, B, tile_offset), B)
        allreduce(stats)
        sigma, alpha = step_schedule(t, T, alpha_range, radius_range)
        apply(stats, sigma, alpha)
    return T
