"""Latency of the reference-named drop-ins at the reference's own per-call sizes (one clip's 12 audio rows, one frame),
next to the NumPy oracle on the same host - small calls are launch- and copy-latency-bound, batches are where the GPU pays."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import acoustic_image_generation_b200 as aig
from acoustic_image_generation_b200 import synth, tables
from oracle import acoustic_oracle as oracle      # timing comparison only (tools/, not product code)


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
