"""Two-GPU test of the native count all-reduce (aig_comm_init / aig_allreduce_counts over NCCL) and of the sharded
evaluation driver.  Needs two CUDA devices; on a one-GPU box it is skipped (the gloo tests cover the host logic)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu
THR = [0.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0]


def _worker(rank, world, port, n, out):
    import torch
    import torch.distributed as dist
    import acoustic_image_generation_b200 as aig
    from acoustic_image_generation_b200 import evaluate, sharding, synth
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('gloo', rank=rank, world_size=world)      # only ships the NCCL unique id
    path = aig.AcousticPath(rank)
    path.init_comm()
    a = synth.smooth_images(n, 80)
    b = synth.smooth_images(n, 81)
    b[::2] = a[::2] * np.float32(0.85) + b[::2] * np.float32(0.15)
    lo, hi = sharding.shard_range(n, rank, world)
    ev = evaluate.AcivwEvaluation(path, THR)
    ev.add_batch(a[lo:hi], b[lo:hi])
    path.allreduce_counts(ev.counts)                                   # native NCCL, device tensor, handle's stream
    path.synchronize()
    host = np.array([rank + 1, 10 * (rank + 1)], np.int64)            # host buffers are staged
    path.allreduce_counts(host)
    if rank == 0:
        np.save(out, np.concatenate([ev.counts.cpu().numpy(), host]))
    dist.barrier()
    path.close()
    dist.destroy_process_group()


def test_native_nccl_allreduce_of_counts(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    import torch.multiprocessing as mp
    from acoustic_image_generation_b200 import synth
    from oracle import acoustic_oracle as oracle
    n, world = 37, 2
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    out = str(tmp_path / 'counts.npy')
    mp.spawn(_worker, args=(world, port, n, out), nprocs=world, join=True)
    got = np.load(out)
    a = synth.smooth_images(n, 80)
    b = synth.smooth_images(n, 81)
    b[::2] = a[::2] * np.float32(0.85) + b[::2] * np.float32(0.15)
    ea, _ = oracle.energy_stage(a, normalize_first=False)
    eb, _ = oracle.energy_stage(b, normalize_first=False)
    _, _, pos, num = oracle.acivw_sweep(ea, eb, THR)
    assert got[:11].tolist() == pos.tolist() and got[11] == num == n
    assert got[12:].tolist() == [3, 30]


def test_allreduce_without_communicator_is_identity():
    import torch
    import acoustic_image_generation_b200 as aig
    path = aig.AcousticPath(0)
    c = torch.arange(12, dtype=torch.int64, device='cuda')
    path.allreduce_counts(c)
    assert c.cpu().tolist() == list(range(12))
    h = np.arange(5, dtype=np.int64)
    path.allreduce_counts(h)
    assert h.tolist() == [0, 1, 2, 3, 4]
    path.close()
