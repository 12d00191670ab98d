"""Latency of the reference-named drop-ins at the reference's own per-call sizes (one clip's 12 audio rows, one frame),
next to the NumPy oracle on the same host - small calls are launch- and copy-latency-bound, batches are where the GPU pays."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import acoustic_image_generation_b200 as aig
from acoustic_image_generation_b200 import synth, tables
from oracle import acoustic_oracle as oracle      # timing comparison (tests/ may use the oracle)


def bench(fn, reps=200):
    for _ in range(10):
        fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e6


bank, dct, lifter, mfnorm = tables.reference_tables()
rows12 = synth.power_frames(1, 0, 'chi2').reshape(-1, 512)[:12].copy()
frame = synth.power_frames(1, 1, 'chi2').reshape(-1, 512)
img = synth.sigmoid_images(1, 0)[0]
audio = synth.audio_rows(12, 0, np.int32)
cases = [
    ('get_feats, 12 rows (one clip of audio)', lambda: aig.get_feats(512, rows12, 12, dct, mfnorm, lifter, bank),
     lambda: oracle.get_feats(512, rows12, 12, dct, mfnorm, lifter, bank)),
    ('get_feats, 1728 rows (one acoustic frame)', lambda: aig.get_feats(512, frame, 12, dct, mfnorm, lifter, bank),
     lambda: oracle.get_feats(512, frame, 12, dct, mfnorm, lifter, bank)),
    ('find_logen, one frame', lambda: aig.find_logen(img.copy()), lambda: oracle.find_logen(img.copy())),
    ('_build_spectrograms_function, 12 x 1024 samples', lambda: aig._build_spectrograms_function(audio),
     lambda: oracle.build_spectrograms(audio)),
]
for name, gpu, cpu in cases:
    print('%-52s GPU drop-in %8.1f us   NumPy oracle %8.1f us' % (name, bench(gpu), bench(cpu, 50)))

# where a small call's time goes: the bare C entry point on preallocated arrays vs the Python wrappers around it
path = aig.default_path()
out12 = np.empty((12, 12), np.float32)
lib, h = path._lib, path._h
print('%-52s %8.1f us' % ('aig_mfcc (ctypes, 12 rows, preallocated, pageable)',
                          bench(lambda: lib.aig_mfcc(h, rows12.ctypes.data, 12, out12.ctypes.data, 0, 1))))
print('%-52s %8.1f us' % ('AcousticPath.mfcc_rows (12 rows)', bench(lambda: path.mfcc_rows(rows12))))
print('%-52s %8.1f us' % ('AcousticPath.set_tables (unchanged tables)', bench(lambda: path.set_tables(bank, dct, lifter, mfnorm))))
import torch
d_rows = torch.from_numpy(rows12).cuda(); d_out = torch.empty((12, 12), device='cuda')
def dev_call():
    lib.aig_mfcc(h, d_rows.data_ptr(), 12, d_out.data_ptr(), 0, 1)
    torch.cuda.synchronize()
print('%-52s %8.1f us' % ('aig_mfcc (device in/out) + synchronize', bench(dev_call)))
