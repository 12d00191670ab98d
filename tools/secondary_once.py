"""Two launches of every secondary kernel on resident inputs (2048 frames) - the program profiled for
profiles/r01_ncu_secondary_kernels.csv:
    ncu --set full --clock-control none --import-source on -k regex:'energy_kernel|heatmap|overlay|ciou|resize_mask|triplet|tile|normalize' \\
        -o gpurun_out/secondary python tools/secondary_once.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import acoustic_image_generation_b200 as aig
from acoustic_image_generation_b200 import synth

p = aig.AcousticPath(0)
n = 2048
img = torch.rand(n, 36, 48, 12, device='cuda')
other = torch.rand(n, 36, 48, 12, device='cuda')
boxes = [torch.from_numpy(np.ascontiguousarray(b)).cuda() for b in synth.flickr_boxes(n, 0)]
frames = torch.randint(0, 256, (n, 224, 298, 3), device='cuda', dtype=torch.uint8)
thr = np.linspace(0, 1, 101)
for _ in range(2):
    energy, mask = p.energy(img, normalize_first=True)
    p.normalize_images(img)
    heat = p.heatmap(energy, 224, 298)
    p.resize_mask(mask, 224, 298)
    p.overlay(heat, frames)
    p.ciou_sweep(mask, *boxes, thr)
    p.tile_mfcc(img[:, 0, 0, :].contiguous())
    p.split_triplets(img)
    p.triplet_mse(img, other)
torch.cuda.synchronize()
print('ok')
