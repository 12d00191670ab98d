import torch, time
n = 1 << 30
src = torch.empty(n, dtype=torch.uint8, pin_memory=True)
dst = torch.empty(n, dtype=torch.uint8, device='cuda')
for k in (1, 2, 4):
    streams = [torch.cuda.Stream() for _ in range(k)]
    part = n // k
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                dst[i * part:(i + 1) * part].copy_(src[i * part:(i + 1) * part], non_blocking=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(k, 'streams: %.1f GB/s' % (n / dt / 1e9))
# chunk size effect on one stream
for mb in (1, 4, 16, 64, 256):
    c = mb << 20
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for off in range(0, n, c):
        dst[off:off + c].copy_(src[off:off + c], non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print('%3d MiB chunks: %.1f GB/s' % (mb, n / dt / 1e9))
