"""Two launches each of aig_energy and aig_acivw_batch on 4096 resident frames (min-max normalisation on), in both forms
(energy_wide = 1: stage2_wide_kernel, 0: stage2_kernel) - the program profiled for profiles/r02_ncu_stage2_wide.csv:
    ncu --set full --clock-control none --import-source on -k regex:stage2 -s 4 -o gpurun_out/wide python tools/energy_once.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import acoustic_image_generation_b200 as aig

p = aig.AcousticPath(0)
n = 4096
img = torch.randn(n, 36, 48, 12, device='cuda') * 12 - 8
other = torch.randn(n, 36, 48, 12, device='cuda') * 12 - 8
thr = torch.tensor(aig.REFERENCE_THRESHOLDS, device='cuda', dtype=torch.float64)
cnt = torch.zeros(12, device='cuda', dtype=torch.int64)
energy = torch.empty(n, 36, 48, device='cuda', dtype=torch.float64)
mask = torch.empty(n, 36, 48, device='cuda', dtype=torch.uint8)
lib, h = p._lib, p._h
for wide in (1, 0, 1, 0):
    p.set_option('energy_wide', wide)
    p._check(lib.aig_energy(h, img.data_ptr(), n, 1, None, energy.data_ptr(), mask.data_ptr(), None))
    p._check(lib.aig_acivw_batch(h, img.data_ptr(), other.data_ptr(), n, 1, thr.data_ptr(), 11, None, None, cnt.data_ptr(), cnt[11:].data_ptr(),
                                 None, None, None, None))
torch.cuda.synchronize()
print('ok')
