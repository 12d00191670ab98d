"""aig_resize_mask / aig_ciou_sweep at the reference's output sizes: packed kernels (mask_packed_kernel.cuh) against the
generic ones, resident inputs.    python tools/mask_probe.py [frames]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import acoustic_image_generation_b200 as aig
from acoustic_image_generation_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
p = aig.AcousticPath(0)
lib, h = p._lib, p._h
imgs = synth.smooth_images(256, 5)
e = imgs.sum(-1)
base = torch.from_numpy((e > e.mean(axis=(1, 2), keepdims=True)).astype(np.uint8)).cuda()
mask = base.repeat((n + 255) // 256, 1, 1)[:n].contiguous()
thr = torch.linspace(0, 1, 101, dtype=torch.float64, device='cuda')
pos = torch.zeros(101, dtype=torch.int64, device='cuda')
num = torch.zeros(1, dtype=torch.int64, device='cuda')
i2 = torch.empty(n, dtype=torch.int64, device='cuda')
u2 = torch.empty(n, dtype=torch.int64, device='cuda')


def timed(fn, reps=9):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


for hh, ww in ((224, 298), (224, 224)):
    up = torch.empty(n, hh, ww, dtype=torch.uint8, device='cuda')
    boxes = [torch.from_numpy(np.ascontiguousarray(b)).cuda() for b in synth.flickr_boxes(n, 0, hh, ww)]
    for packed in (1, 0):
        p.set_option('mask_packed', packed)
        ms = timed(lambda: lib.aig_resize_mask(h, mask.data_ptr(), n, hh, ww, up.data_ptr()))
        print('aig_resize_mask %dx%d mask_packed=%d  %7.3f ms  %6.2f M frames/s  %5.2f TB/s written' %
              (hh, ww, packed, ms, n / ms / 1e3, n * hh * ww / ms / 1e9), flush=True)
        ms = timed(lambda: lib.aig_ciou_sweep(h, mask.data_ptr(), boxes[0].data_ptr(), boxes[1].data_ptr(), boxes[2].data_ptr(),
                                              boxes[3].data_ptr(), n, hh, ww, thr.data_ptr(), 101, i2.data_ptr(), u2.data_ptr(),
                                              pos.data_ptr(), num.data_ptr()))
        print('aig_ciou_sweep  %dx%d mask_packed=%d  %7.3f ms  %6.2f M frames/s (101 thresholds)' % (hh, ww, packed, ms, n / ms / 1e3), flush=True)
    p.set_option('mask_packed', 1)
