import os, sys
sys.path.insert(0, os.getcwd())
import torch
import acoustic_image_generation_b200 as aig
p = aig.AcousticPath(0)
lib, h = p._lib, p._h
def timed(fn, reps=11):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
thr = torch.tensor(aig.REFERENCE_THRESHOLDS, device='cuda', dtype=torch.float64)
cnt = torch.zeros(12, device='cuda', dtype=torch.int64)
for n in (148, 300, 600, 1184, 2048, 4096, 5000, 8192, 16384):
    img = torch.randn(n, 36, 48, 12, device='cuda') * 12 - 8
    other = torch.randn(n, 36, 48, 12, device='cuda') * 12 - 8
    energy = torch.empty(n, 36, 48, device='cuda', dtype=torch.float64)
    mask = torch.empty(n, 36, 48, device='cuda', dtype=torch.uint8)
    out = []
    for wide in (1, 0):
        p.set_option('energy_wide', wide)
        e = timed(lambda: lib.aig_energy(h, img.data_ptr(), n, 1, None, energy.data_ptr(), mask.data_ptr(), None))
        a = timed(lambda: lib.aig_acivw_batch(h, img.data_ptr(), other.data_ptr(), n, 1, thr.data_ptr(), 11, None, None, cnt.data_ptr(), cnt[11:].data_ptr(), None, None, None, None))
        out += [n / e / 1e3, n / a / 1e3]
    p.set_option('energy_wide', 1)
    print('%6d frames: aig_energy(norm) wide %.2f / narrow %.2f M frames/s; aig_acivw_batch(norm) wide %.2f / narrow %.2f M pairs/s' % (n, out[0], out[2], out[1], out[3]), flush=True)
