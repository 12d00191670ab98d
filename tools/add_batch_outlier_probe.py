"""Per-call times of the evaluation step from NumPy arrays (mean / median / max), quiet and in bench.py's situation
(after large staged uploads and a multi-threaded BLAS product, as the NumPy checker leaves the host).  The committed
output (profiles/r02_add_batch_outlier_probe.txt) was made by a first version that also ran bench.py's config_c1 in
the same process; that part was removed because tools/ must not execute the oracle.  Run on a GPU box: python tools/add_batch_outlier_probe.py"""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
import torch
import acoustic_image_generation_b200 as aig
from acoustic_image_generation_b200 import evaluate, synth

p = aig.AcousticPath(0)
ev = evaluate.AcivwEvaluation(p)
real, recon = synth.sigmoid_images(16, 1), synth.sigmoid_images(16, 2)


def series(fn, warm, reps, label):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append((time.perf_counter() - t0) * 1e6)
    a = np.array(ts)
    worst = np.argsort(a)[-3:][::-1]
    print('%-58s mean %7.0f  median %6.0f  max %8.0f us  (worst calls: %s)' % (label, a.mean(), np.median(a), a.max(),
          ', '.join('#%d %.0f' % (i, a[i]) for i in worst)))


for min_bytes in (8 << 20, 1 << 20):
    p.set_option('staged_min_bytes', min_bytes)
    tag = 'staged_min_bytes %d: ' % min_bytes
    series(lambda: ev.add_batch(real, recon), 3, 50, tag + 'quiet, 3 warm-ups')
    time.sleep(0.5)
    series(lambda: ev.add_batch(real, recon), 0, 50, tag + 'after 0.5 s idle, no warm-up')
    power = synth.power_frames(16, 0, 'chi2')
    for _ in range(3):
        p.mfcc_energy(power, flip=True, normalize_first=True)
    series(lambda: ev.add_batch(real, recon), 3, 50, tag + 'after 57 MB staged uploads')
    blas = np.random.default_rng(0).random((2500, 2500))
    blas @ blas
    series(lambda: ev.add_batch(real, recon), 3, 50, tag + 'after a BLAS product')
    d_real = torch.from_numpy(real).cuda()
    series(lambda: ev.add_batch(real, recon), 3, 50, tag + 'after a torch upload')
    series(lambda: p.energy(real), 3, 50, tag + 'find_logen of the 16 frames (1.3 MB in, 0.25 MB out)')

# ---- the worker path for small uploads (staged_solo_bytes 0)
print('--- staged_solo_bytes 0 (small uploads handed to the copy threads) ---')
p.set_option('staged_solo_bytes', 0)
p.set_option('staged_min_bytes', 1 << 20)
series(lambda: ev.add_batch(real, recon), 3, 50, 'workers: quiet, 3 warm-ups')
time.sleep(0.5)
series(lambda: ev.add_batch(real, recon), 0, 50, 'workers: after 0.5 s idle, no warm-up')
blas @ blas
series(lambda: ev.add_batch(real, recon), 3, 50, 'workers: after a BLAS product')
big = np.concatenate([synth.power_frames(64, 1, 'chi2')] * 4, 0)
p.mfcc_energy(big, flip=True, normalize_first=True)
del big
series(lambda: ev.add_batch(real, recon), 3, 50, 'workers: after a 906 MB staged upload')
