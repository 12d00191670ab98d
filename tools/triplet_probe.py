"""Throughput of the N2 channel-triplet kernels on resident device tensors (HBM-bound; prints GB/s)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, time, numpy as np
import acoustic_image_generation_b200 as aig
p = aig.AcousticPath(0)
n = 40000
a = torch.rand(n, 36, 48, 12, device='cuda'); b = torch.rand(n, 36, 48, 12, device='cuda')
for name, fn, bytes_ in (('split', lambda: p.split_triplets(a), 2 * a.numel() * 4), ('mse', lambda: p.triplet_mse(a, b), 2 * a.numel() * 4)):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(name, '%.3f ms  %.0f GB/s' % (ms, bytes_ / ms / 1e6))
