"""Device-level SOM operators of the Pixie hot path, on torch CUDA tensors, plus the
pyFlowSOM-shaped function API (``som``, ``map_data_to_nodes``) the reference binds at
``/root/reference/src/ark/phenotyping/cluster_helpers.py:14`` and calls at ``:106-109`` / ``:152-157``.

torch is plumbing here (device memory, streams, NCCL); every kernel is in libpixie_b200.so and is
reached through the C ABI of ``include/pixie_b200.h``.  There is no CPU fallback: without a CUDA
device or without the built library these functions raise.
"""
import ctypes
import os
import threading

import numpy as np
import torch

from . import _native
from ._native import (FLAG_AUTO, FLAG_FORCE_EXACT, FLAG_FORCE_TC, NSTATS, TILE,  # noqa: F401
                      PixieError)

__all__ = [
    "to_device_matrix", "bmu", "bmu_dists", "cluster_sums", "label_sums", "som_accum", "som_apply", "train_som",
    "som", "map_data_to_nodes", "default_radius", "init_codebook_indices", "default_batches",
    "grid_chebyshev",
]


# ------------------------------------------------------------------------------------------------
# host-side pieces of the algorithm definition
# ------------------------------------------------------------------------------------------------
def grid_chebyshev(xdim, ydim):
    """K x K Chebyshev distance between SOM grid nodes, node k <-> (k // ydim, k % ydim)."""
    k = np.arange(xdim * ydim)
    gx, gy = k // ydim, k % ydim
    return np.maximum(np.abs(gx[:, None] - gx[None, :]),
                      np.abs(gy[:, None] - gy[None, :])).astype(np.float64)


def default_radius(xdim, ydim):
    """Neighbourhood radius range (start, end): pyFlowSOM's default, the 0.67 quantile of the grid
    distances down to 0 (6.0 for a 10x10 map, 11.0 for 20x20)."""
    return float(np.quantile(grid_chebyshev(xdim, ydim), 0.67)), 0.0


def init_codebook_indices(n, K, seed):
    """Seeded choice of K distinct rows for the initial codebook, as the reference dependency
    draws it: numpy's LEGACY generator seeded with ``seed``, ``choice(n, K, replace=False)``
    (pyFlowSOM seeds the global ``np.random``; ``RandomState(seed)`` is the same stream without the
    side effect), so the same seed starts from the same K rows as the reference.  Raises ValueError
    when n < K, like the reference.  Cost: the legacy ``choice`` permutes ``arange(n)`` (8 n bytes,
    ~2 s per 2e8 rows) exactly as it does inside the reference."""
    return np.random.RandomState(seed).choice(n, K, replace=False)


def default_batches(n):
    """Mini-batches per training pass: 32, or the number of 128-row tiles when there are fewer."""
    ntiles = (n + TILE - 1) // TILE
    return max(1, min(32, ntiles))


# ------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------
def _require_cuda(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise PixieError(f"{name} must be a CUDA tensor (no CPU fallback on this path)")


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _check_x(X):
    _require_cuda(X, "X")
    if X.dtype != torch.float32 or X.dim() != 2:
        raise PixieError("X must be a 2-D float32 tensor")
    if X.shape[0] > 1 and X.shape[1] > 1 and X.stride(1) != 1:
        raise PixieError("X rows must be contiguous (stride(1) == 1)")
    n, C = X.shape
    ld = X.stride(0) if n > 1 else max(C, X.stride(0))
    return n, C, ld


def _check_w(W, C):
    _require_cuda(W, "W")
    if W.dtype != torch.float32 or W.dim() != 2 or W.shape[1] != C or not W.is_contiguous():
        raise PixieError("W must be a contiguous [K, C] float32 tensor")
    return W.shape[0]


_ws_cache = {}


def _workspace(n_visit, C, K, device):
    """Per (device, stream, host thread) cached workspace, grown on demand.

    The workspace holds the control block the kernels of one call chain share (codebook norms,
    fix-up counter: memset -> prep -> BMU -> fix-up are separate launches), so two host threads
    must never enqueue on the same one: ``cluster_pixels(multiprocess=True)`` labels FOVs from a
    thread pool and ctypes releases the GIL during the calls."""
    need = _native.lib().pixie_workspace_bytes(int(n_visit), int(C), int(K))
    key = (device.index, torch.cuda.current_stream(device).cuda_stream, threading.get_ident())
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


def to_device_matrix(data, device=None, out=None):
    """Flatten a host [n, C] array (any float dtype) into the device layout the kernels stream:
    fp32, row-major, row pitch rounded up to 4 floats (16 bytes, the TMA stride granule), zero
    padded.  Returns the [n, C] view (stride(0) = pitch)."""
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if isinstance(data, torch.Tensor):
        src = data
    else:
        src = torch.from_numpy(np.ascontiguousarray(data))
    if src.dim() != 2:
        raise PixieError("data must be 2-D")
    n, C = src.shape
    ld = (C + 3) // 4 * 4
    if out is None:
        out = torch.zeros((max(n, 1), ld), dtype=torch.float32, device=device)
    view = out[:n, :C]
    if n:
        view.copy_(src.to(device=device, non_blocking=True))  # fp64 -> fp32 rounds to nearest
    return view


def columns_to_rows(cols, divisor=None, out=None):
    """N2: turn per-channel float64 column buffers into the device matrix in one kernel.

    ``cols`` is a CUDA float64 tensor [C, n] (row c = the Arrow buffer of channel c; stride(1)
    must be 1), ``divisor`` an optional [C] float64 tensor (the normalisation row of
    cluster_helpers.py:244-246).  Returns the [n, C] fp32 view (row pitch a multiple of 4 floats)
    holding ``float32(cols[c, i] / divisor[c])`` -- bit-identical to casting the reference's
    normalised float64 table to fp32."""
    if not isinstance(cols, torch.Tensor) or not cols.is_cuda or cols.dtype != torch.float64 \
            or cols.dim() != 2 or (cols.shape[1] > 1 and cols.stride(1) != 1):
        raise PixieError("cols must be a CUDA float64 tensor [C, n] with contiguous columns")
    C, n = cols.shape
    dev = cols.device
    if divisor is not None:
        divisor = torch.as_tensor(divisor, dtype=torch.float64).to(dev).contiguous()
        if divisor.numel() != C:
            raise PixieError("divisor must hold one value per column")
    ld = (C + 3) // 4 * 4
    if out is None:
        out = torch.zeros((max(n, 1), ld), dtype=torch.float32, device=dev)
    elif out.dtype != torch.float32 or out.dim() != 2 or out.shape[0] < n or out.shape[1] < C \
            or out.stride(1) != 1 or out.device != dev:
        raise PixieError("out must be a CUDA float32 matrix with at least [n, C] elements")
    ld = out.stride(0)
    if n:
        with torch.cuda.device(dev):
            rc = _native.lib().pixie_columns_to_rows_f32(
                _ptr(cols), cols.stride(0) if C > 1 else max(n, cols.stride(0)), n, C,
                _ptr(divisor), _ptr(out), ld, _stream(dev))
        _native.check(rc, "pixie_columns_to_rows_f32")
    return out[:n, :C]


# ------------------------------------------------------------------------------------------------
# operators
# ------------------------------------------------------------------------------------------------
def bmu(X, W, labels=None, want_sums=False, flags=FLAG_AUTO, stats=None):
    """Best-matching-unit labels (1-indexed int32) of every row of X against codebook W.

    Bit-identical to pyFlowSOM.map_data_to_nodes(W, X)[0] evaluated in fp64 on the same
    fp32-representable inputs (cluster_helpers.py:152-157).  With ``want_sums`` also returns the
    per-node channel sums and counts ``SN`` [K, C+1] (float64) of the assigned rows.
    """
    n, C, ld = _check_x(X)
    K = _check_w(W, C)
    dev = X.device
    if labels is None:
        labels = torch.empty(n, dtype=torch.int32, device=dev)
    elif labels.dtype != torch.int32 or labels.numel() != n or not labels.is_contiguous():
        raise PixieError("labels must be a contiguous int32 tensor of length n")
    SN = torch.empty((K, C + 1), dtype=torch.float64, device=dev) if want_sums else None
    if n == 0:
        if SN is not None:
            SN.zero_()
        return (labels, SN) if want_sums else labels
    with torch.cuda.device(dev):
        ws = _workspace(0, C, K, dev)
        rc = _native.lib().pixie_bmu_f32(_ptr(X), n, C, ld, _ptr(W), K, _ptr(labels), _ptr(SN),
                                         _ptr(ws), ws.numel(), flags, _ptr(stats), _stream(dev))
    _native.check(rc, "pixie_bmu_f32")
    return (labels, SN) if want_sums else labels


def bmu_dists(X, W, labels):
    """Exact fp64 distance of each row to its assigned node (map_data_to_nodes()[1])."""
    n, C, ld = _check_x(X)
    K = _check_w(W, C)
    dists = torch.empty(n, dtype=torch.float64, device=X.device)
    if n:
        with torch.cuda.device(X.device):
            rc = _native.lib().pixie_bmu_dist_f64(_ptr(X), n, C, ld, _ptr(W), K, _ptr(labels),
                                                  _ptr(dists), _stream(X.device))
        _native.check(rc, "pixie_bmu_dist_f64")
    return dists


def cluster_sums(X, W, labels=None):
    """(labels, SN): labels plus per-node channel sums / counts -- one fused call."""
    return bmu(X, W, labels=labels, want_sums=True)


def label_sums(X, labels, K):
    """Per-cluster channel sums and counts SN [K, C+1] (float64) for an existing int32 label
    tensor with values in 1..K (the aggregate of pixel_cluster_utils.py:369-374)."""
    n, C, ld = _check_x(X)
    dev = X.device
    _require_cuda(labels, "labels")
    if labels.dtype != torch.int32 or labels.numel() != n or not labels.is_contiguous():
        raise PixieError("labels must be a contiguous int32 tensor of length n")
    SN = torch.zeros((K, C + 1), dtype=torch.float64, device=dev)
    if n:
        with torch.cuda.device(dev):
            ws = _workspace(0, C, K, dev)
            rc = _native.lib().pixie_cluster_sums_f32(_ptr(X), n, C, ld, _ptr(labels), int(K),
                                                      _ptr(SN), _ptr(ws), ws.numel(), _stream(dev))
        _native.check(rc, "pixie_cluster_sums_f32")
    return SN



def _as_i32(t, name, dev=None):
    """Contiguous (hence 256-byte aligned) CUDA int32 copy / view of an integer tensor."""
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(t)
    if dev is None:
        _require_cuda(t, name)
        dev = t.device
    t = t.to(device=dev)
    if t.dtype != torch.int32:
        t = t.to(torch.int32)
    t = t.contiguous()
    if t.data_ptr() % 16:
        t = t.clone()
    return t


def label_histogram(seg_labels, clusters, n_seg, n_clusters, counts=None):
    """N4: counts[s, c] = number of pixels with segmentation label s and cluster c -- the table
    create_c2pc_data gets from groupby(['label', cluster]).size() + pivot (reference
    cell_cluster_utils.py:119-132).  Both inputs are CUDA integer tensors of length n; pairs outside
    [0, n_seg) x [0, n_clusters) (e.g. -1 for "no cluster") are skipped.  Returns (counts int32
    [n_seg, n_clusters], number of skipped pixels); pass ``counts`` to accumulate over chunks."""
    seg = _as_i32(seg_labels, "seg_labels")
    dev = seg.device
    clu = _as_i32(clusters, "clusters", dev)
    if seg.dim() != 1 or clu.shape != seg.shape:
        raise PixieError("seg_labels and clusters must be 1-D tensors of equal length")
    if counts is None:
        counts = torch.zeros((int(n_seg), int(n_clusters)), dtype=torch.int32, device=dev)
    elif counts.dtype != torch.int32 or tuple(counts.shape) != (int(n_seg), int(n_clusters)) \
            or not counts.is_contiguous() or counts.device != dev:
        raise PixieError("counts must be a contiguous CUDA int32 [n_seg, n_clusters] tensor")
    bad = torch.zeros(1, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        rc = _native.lib().pixie_label_histogram_i32(_ptr(seg), _ptr(clu), seg.numel(), int(n_seg),
                                                     int(n_clusters), _ptr(counts), _ptr(bad),
                                                     _stream(dev))
    _native.check(rc, "pixie_label_histogram_i32")
    return counts, bad


def scatter_labels(row_index, column_index, clusters, H, W, id_map=None, unique=False):
    """N4: the [H, W] int16 cluster mask with mask[row_index[i], column_index[i]] =
    id_map[clusters[i]] (reference data_utils.py:523-551).  ``id_map`` is an int16 look-up table
    indexed by cluster value (None = identity).  Duplicate coordinates resolve as numpy's fancy
    assignment does (the last row wins) unless ``unique=True`` promises there are none."""
    r = _as_i32(row_index, "row_index")
    dev = r.device
    c = _as_i32(column_index, "column_index", dev)
    k = _as_i32(clusters, "clusters", dev)
    if r.dim() != 1 or c.shape != r.shape or k.shape != r.shape:
        raise PixieError("row_index, column_index and clusters must be 1-D tensors of equal length")
    lut = None
    if id_map is not None:
        lut = torch.as_tensor(id_map).to(device=dev, dtype=torch.int16).contiguous()
    img = torch.zeros((int(H), int(W)), dtype=torch.int16, device=dev)
    winner = None if unique else torch.empty((int(H), int(W)), dtype=torch.int32, device=dev)
    bad = torch.zeros(1, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        rc = _native.lib().pixie_scatter_labels_i16(
            _ptr(r), _ptr(c), _ptr(k), r.numel(), _ptr(lut), 0 if lut is None else lut.numel(),
            int(H), int(W), _ptr(img), _ptr(winner), _ptr(bad), _stream(dev))
    _native.check(rc, "pixie_scatter_labels_i16")
    return img, bad


def som_accum(X, W32, tile_first, tile_stride, SN=None, flags=FLAG_AUTO, stats=None):
    """One mini-batch of the batch SOM: BMU of the rows of tiles tile_first, tile_first+stride, ...
    against W32 and their per-node sums/counts SN [K, C+1] float64."""
    n, C, ld = _check_x(X)
    K = _check_w(W32, C)
    dev = X.device
    if SN is None:
        SN = torch.empty((K, C + 1), dtype=torch.float64, device=dev)
    ntiles = (n + TILE - 1) // TILE
    nvis = max(0, (ntiles - tile_first + tile_stride - 1) // tile_stride) * TILE
    with torch.cuda.device(dev):
        ws = _workspace(nvis, C, K, dev)
        rc = _native.lib().pixie_som_accum_f32(_ptr(X), n, C, ld, _ptr(W32), K, int(tile_first),
                                               int(tile_stride), _ptr(SN), _ptr(ws), ws.numel(),
                                               flags, _ptr(stats), _stream(dev))
    _native.check(rc, "pixie_som_accum_f32")
    return SN


def som_apply(W64, W32, SN, xdim, ydim, sigma, alpha):
    """Batch update of the fp64 master codebook from SN; refreshes the fp32 copy in place."""
    K, C = W64.shape
    with torch.cuda.device(W64.device):
        rc = _native.lib().pixie_som_apply_f64(_ptr(W64), _ptr(W32), _ptr(SN), xdim, ydim, C,
                                               float(sigma), float(alpha), _stream(W64.device))
    _native.check(rc, "pixie_som_apply_f64")


from .distributed import run_training_steps, step_schedule  # noqa: E402,F401


def train_som(X, W0, xdim, ydim, rlen=1, alpha_range=(0.05, 0.01), radius_range=None,
              batches_per_pass=None, flags=FLAG_AUTO, group=None, tile_offset=0):
    """Batch SOM on a device-resident fp32 matrix.  ``W0`` [K, C] is the initial codebook (any
    float dtype, host or device).  Returns the trained codebook as a float64 CUDA tensor.

    Single GPU (``group is None``): one C call enqueues all rlen * B steps back to back.
    Multi GPU: X is this rank's row shard whose first row is global tile ``tile_offset``; every step
    all-reduces the K x (C+1) statistics over ``group`` (NCCL) between accumulate and apply, so all
    ranks hold the same codebook.
    """
    n, C, ld = _check_x(X)
    dev = X.device
    K = xdim * ydim
    if radius_range is None:
        radius_range = default_radius(xdim, ydim)
    W64 = torch.as_tensor(np.asarray(W0) if not isinstance(W0, torch.Tensor) else W0)
    W64 = W64.to(device=dev, dtype=torch.float64).contiguous().clone()
    if tuple(W64.shape) != (K, C):
        raise PixieError("initial codebook must be [xdim*ydim, C]")
    W32 = torch.empty((K, C), dtype=torch.float32, device=dev)
    SN = torch.zeros((K, C + 1), dtype=torch.float64, device=dev)
    distributed = group is not None
    if batches_per_pass is None:
        if distributed:
            raise PixieError("batches_per_pass must be given explicitly in a multi-GPU run")
        batches_per_pass = default_batches(n)
    B = int(batches_per_pass)
    if not distributed:
        with torch.cuda.device(dev):
            ntiles = (n + TILE - 1) // TILE
            ws = _workspace(((ntiles + B - 1) // B + 1) * TILE, C, K, dev)
            rc = _native.lib().pixie_som_train_f32(
                _ptr(X), n, C, ld, _ptr(W64), _ptr(W32), _ptr(SN), xdim, ydim, int(rlen), B,
                float(alpha_range[0]), float(alpha_range[1]), float(radius_range[0]),
                float(radius_range[1]), _ptr(ws), ws.numel(), flags, _stream(dev))
        _native.check(rc, "pixie_som_train_f32")
        return W64
    import torch.distributed as dist
    # Preferred: the whole run in ONE persistent kernel per rank, the per-step statistics summed
    # across GPUs inside the kernel over NVLink peer memory (symmetric memory mapped by torch).
    # The choice is COLLECTIVE: a rank that spun in the in-kernel handshake while another rank sat
    # in an NCCL all-reduce would hang both, so every rank reports whether it can take the peer
    # path and the minimum decides (the all-reduce doubles as the launch rendezvous).
    peers = _peer_exchange(K, C, dev, group)
    L = _native.lib()
    mine = peers is not None and L.pixie_som_train_peers_supported(
        int(C), int(K), int(ld), int(X.data_ptr() % 16 == 0 or n == 0)) == 1
    vote = torch.tensor([1 if mine else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(vote, op=dist.ReduceOp.MIN, group=group)
    use_peers = bool(int(vote.item()))
    global last_exchange_path
    last_exchange_path = "peer" if use_peers else "nccl"
    if use_peers:
        ptrs, base = peers.claim(int(rlen) * B)
        with torch.cuda.device(dev):
            ws = _workspace(0, C, K, dev)
            arr = (ctypes.c_uint64 * len(ptrs))(*ptrs)
            rc = L.pixie_som_train_peers_f32(
                _ptr(X), n, C, ld, _ptr(W64), _ptr(W32), _ptr(SN), xdim, ydim, int(rlen), B,
                float(alpha_range[0]), float(alpha_range[1]), float(radius_range[0]),
                float(radius_range[1]), int(tile_offset), dist.get_world_size(group),
                dist.get_rank(group), arr, base, _ptr(ws), ws.numel(), flags, _stream(dev))
        # every rank agreed to launch: a refusal here would leave the others waiting -> raise
        _native.check(rc, "pixie_som_train_peers_f32")
        return W64
    # Fallback: one accumulate launch, one NCCL all-reduce and one apply launch per step.
    som_apply(W64, W32, SN, xdim, ydim, 1.0, 0.0)  # W32 = fp32(W64)
    run_training_steps(
        rlen, B, int(tile_offset), alpha_range, radius_range,
        accum=lambda first, stride: som_accum(X, W32, first, stride, SN=SN, flags=flags),
        allreduce=lambda stats: dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group),
        apply=lambda stats, sigma, alpha: som_apply(W64, W32, stats, xdim, ydim, sigma, alpha))
    return W64


# "peer" / "nccl": how the last multi-GPU train_som call exchanged its statistics (bench.py
# records it, the 2-GPU test asserts it)
last_exchange_path = None


class _PeerExchange:
    """Per (group, K, C) symmetric exchange buffers for pixie_som_train_peers_f32."""

    def __init__(self, buf, ptrs):
        self.buf = buf          # keeps the symmetric allocation alive
        self.ptrs = ptrs        # device pointers of every rank's buffer, mapped in this process
        self.flag_base = 0

    def claim(self, nsteps):
        base = self.flag_base
        self.flag_base = (self.flag_base + nsteps + 1) & 0x7FFFFFFF
        return self.ptrs, base


_peer_cache = {}


def _peer_exchange(K, C, dev, group):
    """Symmetric-memory exchange buffers, or None when torch cannot provide peer mappings (the
    NCCL step loop is used then).  Collective: every rank of `group` must call it."""
    import os
    if os.environ.get("PIXIE_DISABLE_PEER", "0") == "1":
        return None
    key = (id(group), K, C, dev.index)
    if key in _peer_cache:
        return _peer_cache[key]
    peers = None
    try:
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        nbytes = _native.lib().pixie_peer_buffer_bytes(int(C), int(K))
        buf = symm_mem.empty((nbytes + 7) // 8, dtype=torch.float64, device=dev)
        buf.zero_()
        hdl = symm_mem.rendezvous(buf, group)
        ptrs = [int(p) for p in hdl.buffer_ptrs]
        torch.cuda.synchronize(dev)
        dist.barrier(group=group)  # every rank's buffer is zeroed before anyone signals into it
        if len(ptrs) == dist.get_world_size(group) and all(ptrs):
            peers = _PeerExchange(buf, ptrs)
    except Exception as exc:  # noqa: BLE001 -- any failure here just selects the NCCL loop
        import warnings
        warnings.warn(f"peer-memory exchange unavailable ({exc!r}); using the NCCL step loop")
        peers = None
    _peer_cache[key] = peers
    return peers


# ------------------------------------------------------------------------------------------------
# pyFlowSOM-shaped function API (the numeric plugin boundary)
# ------------------------------------------------------------------------------------------------
def _default_device():
    if not torch.cuda.is_available():
        raise PixieError("no CUDA device: the Pixie SOM path has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def train_som_online(X, W0, xdim, ydim, rlen=1, alpha_range=(0.05, 0.01), radius_range=None,
                     seed=42):
    """Parity mode: pyFlowSOM's own ONLINE rule (FlowSOM C_SOM) on a device-resident fp32 matrix,
    one sequential CTA (``pixie_som_online_f64``).  The sample sequence is the reference's
    ``srand(seed)`` / ``rand()`` stream.  Returns ``(W64 cuda tensor, iterations executed)``; the
    codebook is bit-identical to ``oracle.som_online`` on the same fp32-representable inputs."""
    n, C, ld = _check_x(X)
    dev = X.device
    K = xdim * ydim
    if n < 1:
        raise PixieError("online training needs at least one row")
    if radius_range is None:
        radius_range = default_radius(xdim, ydim)
    W64 = torch.as_tensor(np.asarray(W0) if not isinstance(W0, torch.Tensor) else W0)
    W64 = W64.to(device=dev, dtype=torch.float64).contiguous().clone()
    if tuple(W64.shape) != (K, C):
        raise PixieError("initial codebook must be [xdim*ydim, C]")
    niter = int(rlen) * n
    idx = np.empty(niter, np.int64)
    L = _native.lib()
    _native.check(L.pixie_libc_sample_indices(int(seed) & 0xFFFFFFFF, n, niter,
                                              idx.ctypes.data_as(ctypes.c_void_p)),
                  "pixie_libc_sample_indices")
    idx_dev = torch.from_numpy(idx).to(dev)
    done = torch.zeros(1, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        rc = L.pixie_som_online_f64(_ptr(X), n, C, ld, _ptr(W64), xdim, ydim, _ptr(idx_dev), niter,
                                    float(alpha_range[0]), float(alpha_range[1]),
                                    float(radius_range[0]), float(radius_range[1]), _ptr(done),
                                    _stream(dev))
    _native.check(rc, "pixie_som_online_f64")
    return W64, int(done.item())


# "batch"  = the B200 production algorithm: mini-batch batch SOM, shards
#            over GPUs, a different algorithm from the reference's (weights are NOT comparable);
# "online" = the reference's own sequential rule on the device, bit-exact to its restatement
#            (one CTA, ~3 us per sample, does not shard);
# "auto"   = online while the run is short enough to be sequential (rlen * n samples at most
#            ONLINE_MAX_ITERS: every table of the reference's tests, small cell tables), batch
#            beyond that (pixel tables, large cell tables).
DEFAULT_ALGORITHM = os.environ.get("PIXIE_SOM_ALGORITHM", "auto")
ONLINE_MAX_ITERS = int(os.environ.get("PIXIE_ONLINE_MAX_ITERS", str(1 << 16)))


def som(data, xdim=10, ydim=10, rlen=10, alpha_range=(0.05, 0.01), radius_range=None, seed=None,
        batches_per_pass=None, device=None, algorithm=None):
    """Drop-in for ``pyFlowSOM.som`` as ark calls it (cluster_helpers.py:106-109): trains an
    xdim x ydim SOM on ``data`` [n, C] and returns the codebook as a float64 ndarray [K, C].

    The initial codebook is the reference's: K rows drawn with numpy's legacy generator seeded
    with ``seed`` (``init_codebook_indices``).  What runs next depends on ``algorithm``
    (default ``DEFAULT_ALGORITHM`` / ``PIXIE_SOM_ALGORITHM`` = "auto"):

    * ``"online"`` -- pyFlowSOM's sequential online rule itself (``train_som_online``): same
      sample stream, same arithmetic, so the same seed gives the reference rule's weights;
    * ``"batch"``  -- the mini-batch BATCH SOM of DESIGN.md section 4.  A different algorithm: the
      neighbourhood is Gaussian, a step moves a node by ``1 - (1 - alpha)^den`` (``den`` = its
      neighbourhood-weighted row count) towards the neighbourhood mean, so with mini-batches of
      ~1e5 rows ``lr_start`` / ``lr_end`` hardly matter and a weights file is NOT reproducible
      against the reference's.  Map quality is what is kept (tests/test_train_gpu.py);
    * ``"auto"``   -- online when ``rlen * n <= ONLINE_MAX_ITERS`` (65,536), else batch."""
    device = torch.device(device) if device is not None else _default_device()
    data = np.asarray(data)
    if data.ndim != 2:
        raise ValueError("data must be a 2-D array")
    n, C = data.shape
    K = xdim * ydim
    idx = init_codebook_indices(n, K, seed)
    X = to_device_matrix(data, device)
    # the initial codebook is taken from the fp32 device matrix so that a caller holding only the
    # device matrix gets the same result
    W0 = X[torch.as_tensor(idx, device=device)].to(torch.float64)
    algorithm = algorithm or DEFAULT_ALGORITHM
    if algorithm == "auto":
        algorithm = "online" if int(rlen) * n <= ONLINE_MAX_ITERS else "batch"
    if algorithm == "online":
        W, _ = train_som_online(X, W0, xdim, ydim, rlen=rlen, alpha_range=alpha_range,
                                radius_range=radius_range, seed=0 if seed is None else seed)
        return W.cpu().numpy()
    if algorithm != "batch":
        raise ValueError("algorithm must be 'auto', 'batch' or 'online'")
    W = train_som(X, W0, xdim, ydim, rlen=rlen, alpha_range=alpha_range,
                  radius_range=radius_range, batches_per_pass=batches_per_pass)
    return W.cpu().numpy()


def map_data_to_nodes(nodes, newdata, device=None, chunk_rows=1 << 20, return_dists=True):
    """Drop-in for ``pyFlowSOM.map_data_to_nodes`` (cluster_helpers.py:152-157): returns
    ``(labels int32 1-indexed, dists float64)`` for host arrays, through the host-buffer C entry
    point (pinned-or-pageable host memory in, chunked H2D / kernel / D2H overlap)."""
    device = torch.device(device) if device is not None else _default_device()
    nodes = np.ascontiguousarray(nodes)
    newdata = np.ascontiguousarray(newdata)
    if nodes.ndim != 2 or newdata.ndim != 2 or nodes.shape[1] != newdata.shape[1]:
        raise ValueError("nodes [K, C] and newdata [m, C] must agree on C")
    m, C = newdata.shape
    K = nodes.shape[0]
    labels = np.empty(m, np.int32)
    dists = np.empty(m, np.float64) if return_dists else None
    if m == 0:
        return labels, dists
    L = _native.lib()
    if newdata.dtype == np.float32 and nodes.dtype == np.float32:
        fn = L.pixie_map_data_to_nodes_host_f32
    else:
        nodes = nodes.astype(np.float64, copy=False)
        newdata = newdata.astype(np.float64, copy=False)
        fn = L.pixie_map_data_to_nodes_host_f64
    rc = fn(nodes.ctypes.data_as(ctypes.c_void_p), K, newdata.ctypes.data_as(ctypes.c_void_p), m, C,
            labels.ctypes.data_as(ctypes.c_void_p),
            dists.ctypes.data_as(ctypes.c_void_p) if return_dists else None,
            device.index if device.index is not None else -1, int(chunk_rows))
    _native.check(rc, "pixie_map_data_to_nodes_host")
    return labels, dists
