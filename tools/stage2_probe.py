"""Quick throughput probe of the stage-2 kernels (energy, ACIVW step, heat maps) on resident inputs:
    python tools/stage2_probe.py [frames]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import acoustic_image_generation_b200 as aig

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
p = aig.AcousticPath(0)
img = torch.randn(n, 36, 48, 12, device='cuda') * 12 - 8            # MFCC-like magnitudes (per-frame minimum far from 0)
other = torch.randn(n, 36, 48, 12, device='cuda') * 12 - 8
unit = torch.rand(n, 36, 48, 12, device='cuda')                    # an already normalised frame: minimum ~ 0
thr = torch.tensor(aig.REFERENCE_THRESHOLDS, device='cuda', dtype=torch.float64)
cnt = torch.zeros(12, device='cuda', dtype=torch.int64)
energy = torch.empty(n, 36, 48, device='cuda', dtype=torch.float64)
mask = torch.empty(n, 36, 48, device='cuda', dtype=torch.uint8)
lib, h = p._lib, p._h


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


def line(name, ms, frames, bytes_per_frame=None):
    extra = '' if bytes_per_frame is None else '  %.2f TB/s' % (frames * bytes_per_frame / ms / 1e9)
    print('%-52s %8.3f ms  %6.2f M frames/s%s' % (name, ms, frames / ms / 1e3, extra), flush=True)


for norm in (0, 1):
    ms = timed(lambda: lib.aig_energy(h, img.data_ptr(), n, norm, None, energy.data_ptr(), mask.data_ptr(), None))
    line('aig_energy normalize_first=%d (%d frames)' % (norm, n), ms, n)
ms = timed(lambda: lib.aig_energy(h, unit.data_ptr(), n, 1, None, energy.data_ptr(), mask.data_ptr(), None))
line('aig_energy normalize_first=1, frame minimum ~ 0 (%d frames)' % n, ms, n)
ms = timed(lambda: lib.aig_energy(h, unit.data_ptr(), n, 0, None, energy.data_ptr(), mask.data_ptr(), None))
line('aig_energy normalize_first=0, values in [0, 1) (%d frames)' % n, ms, n)
del unit
ms = timed(lambda: lib.aig_acivw_batch(h, img.data_ptr(), other.data_ptr(), n, 0, thr.data_ptr(), 11, None, None,
                                       cnt.data_ptr(), cnt[11:].data_ptr(), None, None, None, None))
line('aig_acivw_batch (%d pairs = %d energy maps)' % (n, 2 * n), ms, 2 * n)
for shape in ((224, 298), (224, 224), (112, 150)):
    heat = torch.empty((n,) + shape, device='cuda')
    for count in (2048, n):
        ms = timed(lambda: lib.aig_heatmap(h, energy.data_ptr(), count, shape[0], shape[1], heat.data_ptr()))
        line('aig_heatmap %dx%d bulk copies (%d frames)' % (shape + (count,)), ms, count, shape[0] * shape[1] * 4)
    p.set_option('heat_bulk_store', 0)
    ms = timed(lambda: lib.aig_heatmap(h, energy.data_ptr(), n, shape[0], shape[1], heat.data_ptr()))
    p.set_option('heat_bulk_store', 1)
    line('aig_heatmap %dx%d round-1 kernel (%d frames)' % (shape + (n,)), ms, n, shape[0] * shape[1] * 4)
    ms = timed(lambda: lib.aig_energy_heatmap(h, img.data_ptr(), n, 1, None, None, heat.data_ptr(), shape[0], shape[1]))
    line('aig_energy_heatmap %dx%d one launch (%d frames)' % (shape + (n,)), ms, n, shape[0] * shape[1] * 4)
    del heat
power = torch.rand(4096, 36, 48, 512, device='cuda')
mfcc = torch.empty(4096, 36, 48, 12, device='cuda')
heat = torch.empty(4096, 224, 298, device='cuda')
ms = timed(lambda: lib.aig_mfcc_energy(h, power.data_ptr(), 4096, 1, 1, mfcc.data_ptr(), energy.data_ptr(), mask.data_ptr(), None))
line('aig_mfcc_energy (4096 frames)', ms, 4096, 3628800)
ms = timed(lambda: lib.aig_mfcc_energy_heatmap(h, power.data_ptr(), 4096, 1, 1, mfcc.data_ptr(), energy.data_ptr(), mask.data_ptr(), None,
                                               heat.data_ptr(), 224, 298))
line('aig_mfcc_energy_heatmap 224x298, one persistent kernel', ms, 4096, 3628800 + 224 * 298 * 4)
ms0 = timed(lambda: (lib.aig_mfcc_energy(h, power.data_ptr(), 4096, 1, 1, mfcc.data_ptr(), energy.data_ptr(), mask.data_ptr(), None),
                     lib.aig_heatmap(h, energy.data_ptr(), 4096, 224, 298, heat.data_ptr())))
line('aig_mfcc_energy + aig_heatmap 224x298, two launches', ms0, 4096, 3628800 + 224 * 298 * 4)
del power, mfcc, heat
for small in (1, 2, 16, 64, 128):
    ms = timed(lambda: lib.aig_energy(h, img.data_ptr(), small, 0, None, energy.data_ptr(), mask.data_ptr(), None), 21)
    line('aig_energy cluster form, %d frames (incl. launch)' % small, ms, small)
    ms = timed(lambda: lib.aig_acivw_batch(h, img.data_ptr(), other.data_ptr(), small, 0, thr.data_ptr(), 11, None, None,
                                           cnt.data_ptr(), cnt[11:].data_ptr(), None, None, None, None), 21)
    line('aig_acivw_batch cluster form, %d pairs (incl. launch)' % small, ms, small)
