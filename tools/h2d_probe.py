"""How fast can pinned host memory reach the device on this box?  One vs two copy streams, several chunk sizes."""
import time
import torch

total = 1 << 30
host = torch.empty(total, dtype=torch.uint8, pin_memory=True)
dev = torch.empty(total, dtype=torch.uint8, device='cuda')
for n_streams in (1, 2, 4):
    streams = [torch.cuda.Stream() for _ in range(n_streams)]
    for chunk_mb in (8, 32, 64, 256, 1024):
        chunk = chunk_mb << 20
        best = 1e9
        for it in range(4):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i, off in enumerate(range(0, total, chunk)):
                with torch.cuda.stream(streams[i % n_streams]):
                    dev[off:off + chunk].copy_(host[off:off + chunk], non_blocking=True)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        print('%d stream(s), %4d MiB chunks: %.1f GB/s' % (n_streams, chunk_mb, total / best / 1e9), flush=True)
out = torch.empty(256 << 20, dtype=torch.uint8, pin_memory=True)
s2 = torch.cuda.Stream()
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    dev.copy_(host, non_blocking=True)
    with torch.cuda.stream(s2):
        out.copy_(dev[:256 << 20], non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
print('H2D 1 GiB with concurrent D2H 256 MiB: %.1f GB/s H2D-equivalent' % (total / dt / 1e9))
