"""aig_normalize_images by frame count:  python tools/normalize_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import acoustic_image_generation_b200 as aig

p = aig.AcousticPath(0)
for n in (2048, 4096, 16384, 65536):
    img = torch.randn(n, 36, 48, 12, device='cuda') * 12 - 8
    out = torch.empty_like(img)
    ts = []
    for i in range(12):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        p._check(p._lib.aig_normalize_images(p._h, img.data_ptr(), n, out.data_ptr()))
        b.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(a.elapsed_time(b))
    ms = sorted(ts)[len(ts) // 2]
    print('aig_normalize_images %6d frames  %.3f ms  %.1f M frames/s  %.2f TB/s (read + write)' % (n, ms, n / ms / 1e3, n * 165888 / ms / 1e9), flush=True)
    del img, out
