"""End-to-end frames/s of AcousticPath.mfcc_energy from ordinary (pageable) NumPy arrays - what a tf.py_func caller
hands over - against pinned ones, for BASELINE configs[0]-sized calls (16 frames) and larger batches."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import acoustic_image_generation_b200 as aig
from acoustic_image_generation_b200 import synth

p = aig.AcousticPath(0)
for n in (1, 16, 64, 256):
    pageable = synth.power_frames(n, 0, 'chi2')
    pinned_t = torch.from_numpy(pageable).pin_memory()
    pinned = pinned_t.numpy()
    for name, arr in (('pageable', pageable), ('pinned', pinned)):
        for _ in range(3):
            p.mfcc_energy(arr, flip=True, normalize_first=True)
        reps = max(3, 2000 // max(n, 8))
        t0 = time.perf_counter()
        for _ in range(reps):
            p.mfcc_energy(arr, flip=True, normalize_first=True)
        dt = (time.perf_counter() - t0) / reps
        print('%4d frames %-8s %8.3f ms/call  %8.0f frames/s  %6.1f GB/s' % (n, name, dt * 1e3, n / dt, n * 3538944 / dt / 1e9))
