"""Deterministic synthetic inputs shaped like the reference's datasets.

Host-side NumPy only; shared by the tests, ``bench.py`` and ``oracle/make_golden.py``
so that every party sees the same seeded arrays.  Nothing here touches the GPU or
the oracle.

* ``power_frames``   - [n, 36, 48, 512] float32 per-pixel frequency power, the
  input of ``get_feats`` (dataloader/outdoor_data_mfcc.py:803-804 squares an rFFT
  magnitude, hence non-negative).
* ``sigmoid_images`` - [n, 36, 48, 12] float32 in (0, 1), the range of the UNet's
  sigmoid output (models/unet_acresnet.py:89-94) that ``find_logen`` is fed with.
* ``flickr_boxes``   - FlickrSoundNet-style annotator boxes in 298x224 pixel
  coordinates (dataloader/frames.py:290-299, convert_data2.py:225-260).
"""
from __future__ import annotations

import hashlib

import numpy as np

FRAME_H, FRAME_W, FFT_LEN, MFCC_NUM = 36, 48, 512, 12
POWER_KINDS = ('chi2', 'lognormal', 'floor')        # structured_power_frames() adds spatial structure to 'chi2'


def power_frames(n, seed=0, kind='chi2'):
    """n synthetic multispectral acoustic frames, float32 [n, 36, 48, 512].

    ``chi2``      squared standard normal (a squared rFFT magnitude of noise);
    ``lognormal`` exp(3 * N(0,1)), nine decades of dynamic range;
    ``floor``     1e-6 * U(0,1): every mel band falls under the 0.001 floor.
    """
    rng = np.random.default_rng(seed)
    shape = (n, FRAME_H, FRAME_W, FFT_LEN)
    if kind == 'chi2':
        x = rng.standard_normal(shape, dtype=np.float32)
        return x * x
    if kind == 'lognormal':
        return np.exp(np.float32(3.0) * rng.standard_normal(shape, dtype=np.float32))
    if kind == 'floor':
        return np.float32(1e-6) * rng.random(shape, dtype=np.float32)
    raise ValueError('unknown kind %r (expected one of %s)' % (kind, POWER_KINDS))


def structured_power_frames(n, seed=0, layout_seed=1000, shift=0.0, gain=1000.0):
    """n multispectral frames with spatial structure: squared-normal noise (as ``power_frames('chi2')``) whose spectrum is
    tilted, pixel by pixel, by a field of one to three Gaussian "sources" per frame - the sources boost a band of bins,
    so the 12-channel MFCC image, and with it the energy map and its mean mask, show connected blobs like a real
    acoustic image instead of salt and pepper.  (find_logen's map of a min-max-normalised MFCC image is flat to first
    order - the cepstral basis annihilates constants - so the sources need a strong gain, default 1000 in their band,
    to stand out of the noise; the background stays salt and pepper.)

    The source layout of frame i comes from ``layout_seed`` alone; ``seed`` drives the noise and, scaled by ``shift``
    (pixels, standard deviation), a per-frame displacement of the whole layout.  Two streams with the same
    ``layout_seed`` therefore show the same sources, displaced: their mask IoU spreads over (0, 1) instead of being all
    or nothing, which is what a "real vs reconstructed" comparison (iouenergythreshold.py:213-229) looks like."""
    rng = np.random.default_rng(seed)
    lay = np.random.default_rng(layout_seed)
    x = rng.standard_normal((n, FRAME_H, FRAME_W, FFT_LEN), dtype=np.float32)
    power = x * x
    yy, xx = np.mgrid[0:FRAME_H, 0:FRAME_W].astype(np.float32)
    bins = np.arange(FFT_LEN, dtype=np.float32)
    for i in range(n):
        count = int(lay.integers(1, 4))
        cy, cx = lay.uniform(4, FRAME_H - 4, count), lay.uniform(4, FRAME_W - 4, count)
        sig = lay.uniform(2.5, 7.0, count)
        tone, width = lay.uniform(60.0, 420.0), lay.uniform(30.0, 90.0)
        dy, dx = (rng.normal(0.0, 1.0, 2) * shift).astype(np.float32)
        field = np.zeros((FRAME_H, FRAME_W), np.float32)
        for k in range(count):
            field += np.exp(-((yy - cy[k] - dy) ** 2 + (xx - cx[k] - dx) ** 2) / np.float32(2 * sig[k] * sig[k])).astype(np.float32)
        band = np.exp(-((bins - np.float32(tone)) / np.float32(width)) ** 2).astype(np.float32)
        power[i] *= np.float32(1.0) + np.float32(gain) * field[:, :, None] * band[None, None, :]
    return power


def sigmoid_images(n, seed=0):
    """n synthetic 12-channel acoustic images with values in [0, 1), float32 [n, 36, 48, 12]."""
    rng = np.random.default_rng(seed)
    return rng.random((n, FRAME_H, FRAME_W, MFCC_NUM), dtype=np.float32)


def smooth_images(n, seed=0):
    """Blob-like 12-channel images (a few Gaussian sources per frame) so that the energy
    masks are connected regions rather than salt-and-pepper; float32 [n, 36, 48, 12] in (0, 1)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:FRAME_H, 0:FRAME_W].astype(np.float32)
    out = np.empty((n, FRAME_H, FRAME_W, MFCC_NUM), dtype=np.float32)
    for i in range(n):
        field = np.zeros((FRAME_H, FRAME_W), dtype=np.float32)
        for _ in range(int(rng.integers(1, 4))):
            cy, cx = rng.uniform(0, FRAME_H), rng.uniform(0, FRAME_W)
            s = rng.uniform(2.0, 8.0)
            field += np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / np.float32(2 * s * s)).astype(np.float32)
        chan = rng.uniform(0.2, 1.0, MFCC_NUM).astype(np.float32)
        img = field[:, :, None] * chan[None, None, :] + np.float32(0.05) * rng.random(
            (FRAME_H, FRAME_W, MFCC_NUM), dtype=np.float32)
        out[i] = img / (img.max() + np.float32(1e-3))
    return out


def flickr_boxes(n, seed=0, height=224, width=298):
    """(xmin, xmax, ymin, ymax), each int32 [n, 3]; 1-3 annotators per frame, absent ones all-zero.

    Coordinates are inclusive pixel indices with xmax allowed to equal ``width``
    (convert_data2.py:257-258 rounds 256-px annotations up to 298), which the
    consumer clips."""
    rng = np.random.default_rng(seed)
    xmin = np.zeros((n, 3), np.int32); xmax = np.zeros((n, 3), np.int32)
    ymin = np.zeros((n, 3), np.int32); ymax = np.zeros((n, 3), np.int32)
    count = rng.integers(1, 4, n)
    for i in range(n):
        for c in range(int(count[i])):
            xa, xb = np.sort(rng.integers(0, width + 1, 2))
            ya, yb = np.sort(rng.integers(0, height + 1, 2))
            if xb == xa:
                xb = min(xa + 1, width)
            if xb == 0:
                xb = 1
            if yb == ya:
                yb = min(ya + 1, height)
            xmin[i, c], xmax[i, c], ymin[i, c], ymax[i, c] = xa, xb, ya, yb
    return xmin, xmax, ymin, ymax


def audio_rows(n, seed=0, dtype=np.int32, amplitude=20000.0):
    """n rows of 1024 audio samples (one acoustic frame's worth at 12 288 Hz): a few decaying tones plus noise,
    int32 like the tfrecords' `audio/data` (outdoor_data_mfcc.py:326) or float32."""
    rng = np.random.default_rng(seed)
    t = np.arange(1024) / 12288.0
    out = np.zeros((n, 1024))
    for i in range(n):
        for _ in range(int(rng.integers(1, 5))):
            f = rng.uniform(40.0, 5000.0)
            out[i] += rng.uniform(0.1, 1.0) * np.sin(2 * np.pi * f * t + rng.uniform(0, 6.28)) * np.exp(-t * rng.uniform(0, 30))
        out[i] += 0.05 * rng.standard_normal(1024)
    out *= amplitude / np.abs(out).max()
    return np.rint(out).astype(np.int32) if np.dtype(dtype) == np.int32 else out.astype(np.float32)


def digest(array):
    """Short SHA-256 of an array's bytes; golden files store it to detect generator drift."""
    return hashlib.sha256(np.ascontiguousarray(array).tobytes()).hexdigest()[:16]
