"""Two launches of the round-2 stage-2 kernels on resident inputs (4096 frames; the cluster forms on 16) - the program
profiled for profiles/r02_ncu_stage2_kernels.csv:
    ncu --set full --clock-control none --import-source on -k regex:'stage2|heat_stream' -o gpurun_out/stage2 python tools/secondary_r02.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import acoustic_image_generation_b200 as aig

p = aig.AcousticPath(0)
n = int(os.environ.get('AIG_FRAMES', '4096'))
img = torch.randn(n, 36, 48, 12, device='cuda') * 12 - 8            # MFCC-like magnitudes
other = torch.randn(n, 36, 48, 12, device='cuda') * 12 - 8
thr = torch.tensor(aig.REFERENCE_THRESHOLDS, device='cuda', dtype=torch.float64)
cnt = torch.zeros(12, device='cuda', dtype=torch.int64)
heat = torch.empty(2048, 224, 298, device='cuda')
for _ in range(2):
    energy, mask = p.energy(img, normalize_first=True)                                   # stage2_kernel<1>
    p.acivw_batch(img, other, thr, pos=cnt[:-1], num=cnt[-1:])                           # stage2_kernel<2>
    p.energy(img[:16], normalize_first=True)                                             # stage2_cluster_kernel<1>
    p.acivw_batch(img[:16], other[:16], thr, pos=cnt[:-1], num=cnt[-1:])                 # stage2_cluster_kernel<2>
    p.heatmap(energy[:2048], 224, 298)                                                   # heat_stream_kernel<false, 2>
    p.energy_heatmap(img[:2048], True, 224, 298, want_energy=False, want_mask=False, out=heat)   # heat_stream_kernel<true, 2>
torch.cuda.synchronize()
print('ok')
