"""The reference-shaped evaluation step from host NumPy arrays (add_batch, B = 16: two arrays of 1.3 MB) and other
mid-size pageable uploads: the driver's own pageable path against the library's staging ring at lower thresholds.
Run on a GPU box: python tools/small_upload_probe.py"""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
import torch
import acoustic_image_generation_b200 as aig
from acoustic_image_generation_b200 import evaluate, synth

p = aig.AcousticPath(0)


def lat(fn, reps=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    ts.sort()
    return ts[len(ts) // 2] * 1e6, ts[0] * 1e6


cases = {}
for b in (2, 16, 64):
    real = synth.sigmoid_images(b, 1); recon = synth.sigmoid_images(b, 2)
    cases[b] = (real, recon)
ev = evaluate.AcivwEvaluation(p)
for min_bytes, piece in ((8 << 20, 256 << 10), (256 << 10, 128 << 10), (256 << 10, 256 << 10), (256 << 10, 512 << 10), (1 << 20, 256 << 10), (1 << 20, 512 << 10)):
    p.set_option('staged_min_bytes', min_bytes)
    p.set_option('staged_small_piece_bytes', piece)
    for streaming in (-1, 0):
        p.set_option('host_copy_streaming', streaming)
        out = []
        for b, (real, recon) in cases.items():
            out.append('B=%d: %.0f us (min %.0f)' % ((b,) + lat(lambda: ev.add_batch(real, recon))))
        e = lat(lambda: p.energy(cases[64][0]), 100)
        print('staged_min_bytes %8d piece %7d streaming %2d | add_batch %s | energy(64 frames, 5.3 MB) %.0f us' % (min_bytes, piece, streaming, ' | '.join(out), e[0]))

# ---- the same step while the host's cores are busy: NumPy's BLAS threads keep spinning for a while after a product (what
# bench.py's config_c1 does just before timing add_batch), and a reference process has TensorFlow's threads besides
print('--- busy host: a 16-thread BLAS product right before every timing loop ---')
p.set_option('staged_min_bytes', 1 << 20)
p.set_option('staged_small_piece_bytes', 512 << 10)
p.set_option('host_copy_streaming', -1)
m = np.random.default_rng(0).random((3000, 3000))
for solo_bytes, min_bytes in ((0, 1 << 20), (4 << 20, 1 << 20), (4 << 20, 8 << 20)):
    p.set_option('staged_solo_bytes', solo_bytes)
    p.set_option('staged_min_bytes', min_bytes)
    out = []
    for b, (real, recon) in cases.items():
        quiet = lat(lambda: ev.add_batch(real, recon))
        m @ m
        busy = lat(lambda: ev.add_batch(real, recon), 50)
        out.append('B=%d: quiet %.0f us, after BLAS %.0f us' % (b, quiet[0], busy[0]))
    m @ m
    e = lat(lambda: p.energy(cases[64][0]), 50)
    print('staged_solo_bytes %8d staged_min_bytes %8d | add_batch %s | energy(64 frames) after BLAS %.0f us' % (solo_bytes, min_bytes, ' | '.join(out), e[0]))
