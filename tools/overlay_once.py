"""Two launches of the overlay kernel on 4096 resident frames (224 x 298) - the program profiled for profiles/r02_ncu_overlay*.csv:
    ncu --set full --clock-control none --import-source on -k regex:overlay -o gpurun_out/overlay python tools/overlay_once.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import acoustic_image_generation_b200 as aig

p = aig.AcousticPath(0)
n = 4096
heat = torch.rand(n, 224, 298, device='cuda')
bgr = torch.randint(0, 256, (n, 224, 298, 3), device='cuda', dtype=torch.uint8)
for mode in (1, 0, 1, 0):
    p.set_option('overlay_luma', mode)
    p.overlay(heat, bgr)
torch.cuda.synchronize()
print('ok')
