"""Host-side construction of the MFCC tables (float64), same formulas and names as the reference.

These are the few-kilobyte constants every reference caller rebuilds on each call
(dataloader/outdoor_data_mfcc.py:806-820); here they are built once and handed to
libaig through aig_set_tables().  They are tables, not the hot path: the per-pixel
arithmetic runs on the GPU only.
"""
from __future__ import annotations

import functools

import numpy as np

FFT_LEN, FILTER_NUM, MFCC_NUM, LIFTER_NUM = 512, 24, 12, 22
LO_FREQ, HI_FREQ = 0, 6400


def createfilters(fft_len, filter_num, lo_freq, hi_freq, samp_freq):
    """Triangular mel filter bank [fft_len, filter_num], float64.

    Drop-in for createfilters (dataloader/outdoor_data_mfcc.py:826-849 and its eight copies):
    filter_num + 2 corner frequencies equally spaced on the HTK mel scale
    mel = 1127 ln(1 + f/700), mapped to FFT bins with floor(f / samp_freq * (fft_len - 1) * 2);
    triangle f rises linearly from corner f to f+1 and falls to corner f+2.
    """
    mel_lo = 1127 * np.log(1 + lo_freq / 700.0)
    mel_hi = 1127 * np.log(1 + hi_freq / 700.0)
    corners_hz = 700.0 * (np.exp(np.linspace(mel_lo, mel_hi, filter_num + 2) / 1127.0) - 1)
    corners = np.floor(corners_hz / float(samp_freq) * (fft_len - 1) * 2).astype('int')
    bank = np.zeros((fft_len, filter_num))
    for f in range(filter_num):
        a, b, c = int(corners[f]), int(corners[f + 1]), int(corners[f + 2])
        bank[a:b + 1, f] = np.linspace(0, 1, b - a + 1)
        bank[b:c + 1, f] = np.linspace(1, 0, c - b + 1)
    return bank


def mfcc_constants(filter_num=FILTER_NUM, mfcc_num=MFCC_NUM, lifter_num=LIFTER_NUM):
    """(dct_base [filter_num, mfcc_num], lifter [mfcc_num], mfnorm), float64
    (outdoor_data_mfcc.py:813-818; find_logen rebuilds the same, iouenergythreshold.py:304-308)."""
    dct_base = np.zeros((filter_num, mfcc_num))
    for m in range(mfcc_num):
        dct_base[:, m] = np.cos((m + 1) * np.pi / filter_num * (np.arange(filter_num) + 0.5))
    lifter = 1 + (lifter_num / 2) * np.sin(np.pi * (1 + np.arange(mfcc_num)) / lifter_num)
    mfnorm = np.sqrt(2.0 / filter_num)
    return dct_base, lifter, mfnorm


@functools.lru_cache(maxsize=4)
def tukey_window(m=1024, alpha=0.75):
    """The window of _build_spectrograms_function: scipy.signal.tukey(1024, alpha=0.75)
    (outdoor_data_mfcc.py:799), float64: raised-cosine tapers over alpha/2 of each end, flat in between."""
    n = np.arange(0, m)
    width = int(np.floor(alpha * (m - 1) / 2.0))
    rise = 0.5 * (1 + np.cos(np.pi * (-1 + 2.0 * n[:width + 1] / alpha / (m - 1))))
    fall = 0.5 * (1 + np.cos(np.pi * (-2.0 / alpha + 1 + 2.0 * n[m - width - 1:] / alpha / (m - 1))))
    win = np.concatenate((rise, np.ones(m - 2 * (width + 1)), fall))
    win.setflags(write=False)
    return win


@functools.lru_cache(maxsize=8)
def butter_lowpass(sample_rate=12288, cutoff=125, order=10):
    """(b, a, zi) of butter_lowpass (outdoor_data_mfcc.py:565-569) plus the lfilter_zi state filtfilt starts from.
    Filter design is table building (a dozen coefficients, via scipy like the reference); the filtering runs on the GPU."""
    from scipy import signal
    b, a = signal.butter(order, cutoff / (0.5 * sample_rate), btype='low', analog=False)
    return np.ascontiguousarray(b), np.ascontiguousarray(a), np.ascontiguousarray(signal.lfilter_zi(b, a))


_JET_SEGMENTS = {   # matplotlib's `jet` LinearSegmentedColormap (x, y0, y1) anchors
    'red': ((0.00, 0, 0), (0.35, 0, 0), (0.66, 1, 1), (0.89, 1, 1), (1.00, 0.5, 0.5)),
    'green': ((0.000, 0, 0), (0.125, 0, 0), (0.375, 1, 1), (0.640, 1, 1), (0.910, 0, 0), (1.000, 0, 0)),
    'blue': ((0.00, 0.5, 0.5), (0.11, 1, 1), (0.34, 1, 1), (0.65, 0, 0), (1.00, 0, 0)),
}


@functools.lru_cache(maxsize=1)
def jet_lut(n=256):
    """plt.cm.jet as a [n, 3] uint8 table (showvideo.py:228 `imshow(map, cmap=plt.cm.jet, alpha=0.7)`): matplotlib's
    piecewise-linear segment data sampled at n points, then `(rgb * 255).astype(uint8)` as Colormap(bytes=True) does."""
    lut = np.zeros((n, 3))
    xind = (n - 1) * np.linspace(0, 1, n)
    for c, name in enumerate(('red', 'green', 'blue')):
        data = np.array(_JET_SEGMENTS[name], dtype=float)
        x, y0, y1 = data[:, 0] * (n - 1), data[:, 1], data[:, 2]
        ind = np.searchsorted(x, xind)[1:-1]
        dist = (xind[1:-1] - x[ind - 1]) / (x[ind] - x[ind - 1])
        lut[:, c] = np.concatenate([[y1[0]], dist * (y0[ind] - y1[ind - 1]) + y1[ind - 1], [y0[-1]]])
    out = (np.clip(lut, 0.0, 1.0) * 255).astype(np.uint8)
    out.setflags(write=False)
    return out


@functools.lru_cache(maxsize=1)
def reference_tables():
    """(filter_mat, dct_base, lifter, mfnorm) of the reference configuration:
    createfilters(512, 24, 0, 6400, 12800) (outdoor_data_mfcc.py:820)."""
    bank = createfilters(FFT_LEN, FILTER_NUM, LO_FREQ, HI_FREQ, 2 * HI_FREQ)
    dct_base, lifter, mfnorm = mfcc_constants()
    for arr in (bank, dct_base, lifter):
        arr.setflags(write=False)
    return bank, dct_base, lifter, float(mfnorm)
