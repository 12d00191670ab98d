"""aig_energy_heatmap at 224 x 298 on 4096 resident MFCC-like frames, warp-specialised forms and the sequential one - the
program profiled for profiles/r02_ncu_energy_heat_ws.csv:
    ncu --set full --clock-control none --import-source on -k regex:'energy_heat_ws|heat_stream' -o gpurun_out/ws python tools/ws_once.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import acoustic_image_generation_b200 as aig

p = aig.AcousticPath(0)
n = 4096
img = torch.randn(n, 36, 48, 12, device='cuda') * 12 - 8
heat = torch.empty(n, 224, 298, device='cuda')
for ws in (1, 0):
    p.set_option('energy_heat_ws', ws)
    for _ in range(2):
        p._check(p._lib.aig_energy_heatmap(p._h, img.data_ptr(), n, 1, None, None, heat.data_ptr(), 224, 298))
torch.cuda.synchronize()
print('ok')
