"""Two launches of the packed mask kernels on 4096 resident frames - the program profiled for profiles/r02_ncu_mask_packed.csv:
    ncu --set full --clock-control none --import-source on -k regex:'packed' -o gpurun_out/mask python tools/mask_once.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import acoustic_image_generation_b200 as aig
from acoustic_image_generation_b200 import synth

p = aig.AcousticPath(0)
n = 4096
imgs = synth.smooth_images(256, 5)
e = imgs.sum(-1)
mask = torch.from_numpy((e > e.mean(axis=(1, 2), keepdims=True)).astype(np.uint8)).cuda().repeat(n // 256, 1, 1).contiguous()
boxes = [torch.from_numpy(np.ascontiguousarray(b)).cuda() for b in synth.flickr_boxes(n, 0)]
thr = np.linspace(0, 1, 101)
for _ in range(2):
    p.resize_mask(mask, 224, 298)
    p.ciou_sweep(mask, *boxes, thr)
torch.cuda.synchronize()
print('ok')
