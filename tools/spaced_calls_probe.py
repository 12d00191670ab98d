"""The 16-frame pageable call (BASELINE configs[0]) back to back and with idle gaps between calls (the staging threads
asleep each time - a caller that does other work between batches), against the driver's own pageable path.
Run on a GPU box: python tools/spaced_calls_probe.py"""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
import torch
import acoustic_image_generation_b200 as aig
from acoustic_image_generation_b200 import synth

p = aig.AcousticPath(0)
power = synth.power_frames(16, 0, 'chi2')
out = (np.empty((16, 36, 48, 12), np.float32), np.empty((16, 36, 48), np.float64), np.empty((16, 36, 48), np.uint8))


def series(gap_s, reps, label):
    for _ in range(3):
        p.mfcc_energy(power, flip=True, normalize_first=True, out=out)
    ts = []
    for _ in range(reps):
        if gap_s:
            time.sleep(gap_s)
        t0 = time.perf_counter(); p.mfcc_energy(power, flip=True, normalize_first=True, out=out); ts.append((time.perf_counter() - t0) * 1e3)
    a = np.array(ts)
    print('%-70s mean %6.3f  median %6.3f  p90 %6.3f  max %6.3f ms' % (label, a.mean(), np.median(a), np.percentile(a, 90), a.max()))


for threads in (-1, 4, 0):
    p.set_option('host_copy_threads', threads)
    tag = 'host_copy_threads %2d%s: ' % (threads, ' (driver pageable path)' if threads == 0 else '')
    series(0, 40, tag + 'back to back')
    series(0.002, 40, tag + '2 ms between calls')
    series(0.02, 40, tag + '20 ms between calls')
    series(0.2, 15, tag + '200 ms between calls')
