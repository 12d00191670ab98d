"""A few seconds' throughput check of the stage-2 kernels (the lines of tools/stage2_probe.py that move when the pixel
arithmetic changes):  python tools/stage2_quick.py [frames]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import acoustic_image_generation_b200 as aig

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
p = aig.AcousticPath(0)
img = torch.randn(n, 36, 48, 12, device='cuda') * 12 - 8
other = torch.randn(n, 36, 48, 12, device='cuda') * 12 - 8
unit = torch.rand(n, 36, 48, 12, device='cuda')
thr = torch.tensor(aig.REFERENCE_THRESHOLDS, device='cuda', dtype=torch.float64)
cnt = torch.zeros(12, device='cuda', dtype=torch.int64)
energy = torch.empty(n, 36, 48, device='cuda', dtype=torch.float64)
mask = torch.empty(n, 36, 48, device='cuda', dtype=torch.uint8)
heat = torch.empty(n, 224, 298, device='cuda')
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


def line(name, ms, frames):
    print('%-58s %8.3f ms  %6.2f M frames/s' % (name, ms, frames / ms / 1e3), flush=True)


line('aig_energy normalize_first=0, MFCC-like', timed(lambda: lib.aig_energy(h, img.data_ptr(), n, 0, None, energy.data_ptr(), mask.data_ptr(), None)), n)
line('aig_energy normalize_first=1, MFCC-like', timed(lambda: lib.aig_energy(h, img.data_ptr(), n, 1, None, energy.data_ptr(), mask.data_ptr(), None)), n)
line('aig_energy normalize_first=1, minimum ~ 0', timed(lambda: lib.aig_energy(h, unit.data_ptr(), n, 1, None, energy.data_ptr(), mask.data_ptr(), None)), n)
line('aig_energy normalize_first=0, values in [0, 1)', timed(lambda: lib.aig_energy(h, unit.data_ptr(), n, 0, None, energy.data_ptr(), mask.data_ptr(), None)), n)
line('aig_energy, energy only (no mask)', timed(lambda: lib.aig_energy(h, unit.data_ptr(), n, 0, None, energy.data_ptr(), None, None)), n)
p.set_option('energy_wide', 0)
line('aig_energy normalize_first=0, MFCC-like, energy_wide=0', timed(lambda: lib.aig_energy(h, img.data_ptr(), n, 0, None, energy.data_ptr(), mask.data_ptr(), None)), n)
line('aig_energy normalize_first=1, MFCC-like, energy_wide=0', timed(lambda: lib.aig_energy(h, img.data_ptr(), n, 1, None, energy.data_ptr(), mask.data_ptr(), None)), n)
p.set_option('energy_wide', 1)
line('aig_acivw_batch (energy maps)', timed(lambda: lib.aig_acivw_batch(h, unit.data_ptr(), other.data_ptr(), n, 0, thr.data_ptr(), 11, None, None,
                                                                       cnt.data_ptr(), cnt[11:].data_ptr(), None, None, None, None)), 2 * n)
for ws in (1, 0):
    p.set_option('energy_heat_ws', ws)
    line('aig_energy_heatmap 224x298 one launch, energy_heat_ws=%d' % ws, timed(lambda: lib.aig_energy_heatmap(h, img.data_ptr(), n, 1, None, None, heat.data_ptr(), 224, 298)), n)
    line('aig_energy_heatmap 224x298 + energy + mask, energy_heat_ws=%d' % ws,
         timed(lambda: lib.aig_energy_heatmap(h, img.data_ptr(), n, 1, energy.data_ptr(), mask.data_ptr(), heat.data_ptr(), 224, 298)), n)
    line('aig_energy_heatmap 224x224 one launch, energy_heat_ws=%d' % ws,
         timed(lambda: lib.aig_energy_heatmap(h, img.data_ptr(), n, 1, None, None, heat.data_ptr(), 224, 224)), n)
p.set_option("energy_heat_ws", 1)
for small in (1, 16):
    line('aig_energy cluster form, %d frames (incl. launch)' % small,
         timed(lambda: lib.aig_energy(h, img.data_ptr(), small, 0, None, energy.data_ptr(), mask.data_ptr(), None), 21), small)
