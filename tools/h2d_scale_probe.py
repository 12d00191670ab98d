"""What the host <-> device fabric of one box gives N GPUs at once (end-to-end scaling 1.00 / 0.99 / 0.52 / 0.41
at N = 1 / 2 / 4 / 8 was unexplained).  Run under torchrun, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/h2d_scale_probe.py

Every rank copies the same 906 MB pinned buffer (256 frames of spectra, bench.py's end-to-end batch) host -> device in
64 MiB cudaMemcpyAsync pieces, five passes, (a) alone on its link, one rank after the other, (b) all ranks at once,
(c) all at once with the result-sized device -> host copies on a second stream, (d) the library call itself
(AcousticPath.mfcc_energy on the same pinned arrays), all at once.  Prints per-rank and whole-box GB/s; no NCCL traffic
inside the timed regions (gloo barriers on the CPU).
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import acoustic_image_generation_b200 as aig

rank, world, local = (int(os.environ.get(k, d)) for k, d in (('RANK', '0'), ('WORLD_SIZE', '1'), ('LOCAL_RANK', '0')))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('gloo')
frames, passes, piece = 256, 5, (64 << 20) // 4
h_in = torch.empty((frames, 36, 48, 512), dtype=torch.float32, pin_memory=True)
h_in.uniform_(0.0, 4.0)
d_in = torch.empty_like(h_in, device=dev)
h_out = [torch.empty(s, dtype=t, pin_memory=True) for s, t in (((frames, 36, 48, 12), torch.float32), ((frames, 36, 48), torch.float64), ((frames, 36, 48), torch.uint8))]
d_out = [torch.empty_like(h, device=dev) for h in h_out]
up, down = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
src, dst = h_in.view(-1), d_in.view(-1)
up_bytes = src.numel() * 4
down_bytes = sum(h.numel() * h.element_size() for h in h_out)


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def copies(with_d2h):
    t0 = time.perf_counter()
    for _ in range(passes):
        with torch.cuda.stream(up):
            for lo in range(0, src.numel(), piece):
                dst[lo:lo + piece].copy_(src[lo:lo + piece], non_blocking=True)
        if with_d2h:
            with torch.cuda.stream(down):
                for h, d in zip(h_out, d_out):
                    h.copy_(d, non_blocking=True)
        up.synchronize()
        down.synchronize()
    return time.perf_counter() - t0


def gather(value):
    if world == 1:
        return [value]
    out = [None] * world
    dist.all_gather_object(out, value)
    return out


def report(name, seconds, nbytes):
    per_rank = gather(passes * nbytes / seconds / 1e9)
    slowest = max(gather(seconds))
    if rank == 0:
        print('%-58s per rank %s GB/s   whole box %.1f GB/s' % (name, ' '.join('%5.1f' % v for v in per_rank), world * passes * nbytes / slowest / 1e9), flush=True)


copies(True)
barrier()
# (a) one rank at a time
alone = 0.0
for r in range(world):
    barrier()
    if r == rank:
        alone = copies(False)
barrier()
per_rank = gather(passes * up_bytes / alone / 1e9)
if rank == 0:
    print('N = %d ranks on %s, %d host CPUs' % (world, torch.cuda.get_device_name(local), os.cpu_count()))
    print('%-58s per rank %s GB/s   (sum %.1f)' % ('(a) H2D alone, one rank at a time', ' '.join('%5.1f' % v for v in per_rank), sum(per_rank)), flush=True)
barrier()
report('(b) H2D, all ranks at once', copies(False), up_bytes)
barrier()
report('(c) H2D + result-sized D2H, all ranks at once', copies(True), up_bytes + down_bytes)
path = aig.AcousticPath(local)
np_in, np_out = h_in.numpy(), tuple(h.numpy() for h in h_out)
for _ in range(2):
    path.mfcc_energy(np_in, flip=True, normalize_first=True, out=np_out)
barrier()
t0 = time.perf_counter()
for _ in range(passes):
    path.mfcc_energy(np_in, flip=True, normalize_first=True, out=np_out)
torch.cuda.synchronize()
report('(d) AcousticPath.mfcc_energy(pinned) = (c) + the kernels', time.perf_counter() - t0, up_bytes + down_bytes)
barrier()
if world > 1:
    dist.destroy_process_group()
