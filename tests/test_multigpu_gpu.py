"""2-GPU NCCL test of sharded training + assignment (skipped with fewer than 2 GPUs).
The 2-rank trained codebook must equal the single-GPU one (same algorithm, same global
mini-batches, one all-reduce per step) and the oracle's within 1e-4 relative."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from conftest import pixie_like

pytestmark = pytest.mark.gpu
XD, YD, B = 10, 10, 8


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, X, W0, out):
    from ark_analysis_b200 import distributed
    from ark_analysis_b200 import som as S
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    lo, hi = distributed.tile_aligned_row_shards(X.shape[0], world)[rank]
    Xd = S.to_device_matrix(X[lo:hi], torch.device("cuda", rank))
    W = S.train_som(Xd, W0, XD, YD, rlen=2, batches_per_pass=B, group=dist.group.WORLD,
                    tile_offset=lo // 128)
    labels = S.bmu(Xd, W.to(torch.float32))
    torch.cuda.synchronize()
    # the fused peer-memory path (not the NCCL fallback) must have been taken
    assert S.last_exchange_path == "peer", "peer-memory exchange not used"
    # a rank WITHOUT rows (fewer tiles than ranks) still takes part in every exchange
    small = X[:100]
    slo, shi = distributed.tile_aligned_row_shards(small.shape[0], world)[rank]
    Xs = S.to_device_matrix(small[slo:shi], torch.device("cuda", rank))
    Ws = S.train_som(Xs, W0[:16], 4, 4, rlen=3, batches_per_pass=1, group=dist.group.WORLD,
                     tile_offset=slo // 128)
    assert S.last_exchange_path == "peer"
    np.save(out % ("s", rank), Ws.cpu().numpy())
    # and the NCCL step loop gives the same codebook
    os.environ["PIXIE_DISABLE_PEER"] = "1"
    S._peer_cache.clear()
    W2 = S.train_som(Xd, W0, XD, YD, rlen=2, batches_per_pass=B, group=dist.group.WORLD,
                     tile_offset=lo // 128)
    assert float((W2 - W).abs().max()) <= 1e-9 * float(W.abs().max())
    np.save(out % ("w", rank), W.cpu().numpy())
    np.save(out % ("l", rank), labels.cpu().numpy())
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_training_and_assignment(tmp_path):
    from ark_analysis_b200 import som as S
    n, C = 128 * 301 + 17, 24
    X = pixie_like(n, C, seed=21)
    idx = oracle.init_codebook_indices(n, XD * YD, 3)
    W0 = X[idx].astype(np.float64)
    out = str(tmp_path / "%s_rank%d.npy")
    mp.spawn(_worker, args=(2, _free_port(), X, W0, out), nprocs=2, join=True)
    w0, w1 = np.load(out % ("w", 0)), np.load(out % ("w", 1))
    np.testing.assert_array_equal(w0, w1)  # identical codebook on every rank
    ref = oracle.som_batch(X, XD, YD, rlen=2, batches_per_pass=B, init_idx=idx)
    assert np.abs(w0 - ref).max() / np.abs(ref).max() < 1e-4
    single = S.train_som(S.to_device_matrix(X), W0, XD, YD, rlen=2, batches_per_pass=B)
    assert np.abs(w0 - single.cpu().numpy()).max() / np.abs(ref).max() < 1e-6
    # the shard-less rank: same codebook on both ranks, equal to the single-GPU one
    s0, s1 = np.load(out % ("s", 0)), np.load(out % ("s", 1))
    np.testing.assert_array_equal(s0, s1)
    alone = S.train_som(S.to_device_matrix(X[:100]), W0[:16], 4, 4, rlen=3, batches_per_pass=1)
    assert np.abs(s0 - alone.cpu().numpy()).max() <= 1e-9 * np.abs(s0).max()
    labels = np.concatenate([np.load(out % ("l", 0)), np.load(out % ("l", 1))])
    want, _ = oracle.map_data_to_nodes_f32(w0.astype(np.float32), X)
    np.testing.assert_array_equal(labels, want)
