"""World-size-2 gloo tests of the N > 1 host logic (runs on CPU): sharding, the count-vector all-reduce
and the AUC on the reduced counts equal the single-process evaluation."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from acoustic_image_generation_b200 import sharding, synth  # noqa: E402
from oracle import acoustic_oracle as oracle  # noqa: E402

THR = list(oracle.REFERENCE_THRESHOLDS)


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 100, 1001):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def test_shard_clips_keeps_clips_whole():
    for fpc in (120, 300):
        spans = [sharding.shard_clips(33, fpc, r, 4) for r in range(4)]
        assert spans[0][2] == 0 and spans[-1][3] == 33 * fpc
        for c0, c1, f0, f1 in spans:
            assert f0 == c0 * fpc and f1 == c1 * fpc


def test_merge_counts():
    assert sharding.merge_counts([[1, 2, 3], [10, 20, 30]]).tolist() == [11, 22, 33]


def _masks(n):
    a = synth.smooth_images(n, 50)
    b = synth.smooth_images(n, 51)
    b[::2] = a[::2] * np.float32(0.8) + b[::2] * np.float32(0.2)
    _, ma = oracle.energy_stage(a, normalize_first=False)
    _, mb = oracle.energy_stage(b, normalize_first=False)
    return ma, mb


def _worker(rank, world, port, n, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    ma, mb = _masks(n)
    lo, hi = sharding.shard_range(n, rank, world)
    scores = [oracle.iou_pair(x, y)[2] for x, y in zip(ma[lo:hi], mb[lo:hi])]
    pos, num = oracle.success_counts(scores, THR)
    counts = torch.from_numpy(np.concatenate([pos, [num]]).astype(np.int64))
    sharding.allreduce_counts(counts)
    arr = np.concatenate([np.zeros(3, np.int64), [rank + 1]]).astype(np.int64)   # NumPy path of the helper
    sharding.allreduce_counts(arr)
    if rank == 0:
        np.save(out, np.concatenate([counts.numpy(), arr]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_allreduce_equals_single_process(tmp_path):
    n, world = 21, 2
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    out = str(tmp_path / 'counts.npy')
    mp.spawn(_worker, args=(world, port, n, out), nprocs=world, join=True)
    got = np.load(out)
    ma, mb = _masks(n)
    scores = [oracle.iou_pair(x, y)[2] for x, y in zip(ma, mb)]
    pos, num = oracle.success_counts(scores, THR)
    assert got[:11].tolist() == pos.tolist() and got[11] == num == n
    assert got[12:].tolist() == [0, 0, 0, 3]
    # the AUC on the reduced counts is the single-process AUC
    assert oracle.auc(THR, oracle.success_rates(got[:11], got[11])) == oracle.auc(THR, oracle.success_rates(pos, num))


def test_allreduce_without_process_group_is_identity():
    c = np.arange(5, dtype=np.int64)
    assert sharding.allreduce_counts(c) is c
