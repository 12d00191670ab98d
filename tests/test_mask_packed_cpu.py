"""The arithmetic of csrc/mask_packed_kernel.cuh restated in NumPy and checked against the oracle on the CPU: reduced tap
weights, two 16-bit pixels per 32-bit multiply-add with the threshold in bit 15, the unclamped periodic vertical taps over
rows padded with one copy of each edge row, and the rectangle-count form of the consensus IoU (showimages_bb.py:288-321).
No GPU: this pins the algorithm the packed kernels implement; tests/test_gpu_round2.py pins the kernels themselves."""
from math import gcd

import numpy as np
import pytest

from acoustic_image_generation_b200 import synth
from oracle import acoustic_oracle as oracle

SRC_H, SRC_W = 36, 48
SIZES = [(224, 298), (224, 224)]


def _geom(h, w):
    fx = gcd(gcd(2 * SRC_W, 2 * w), abs(SRC_W - w))
    fy = gcd(gcd(2 * SRC_H, 2 * h), abs(SRC_H - h))
    return fx, fy, 2 * w // fx, 2 * h // fy


def _masks(n, seed):
    rng = np.random.default_rng(seed)
    e = synth.smooth_images(n, seed).sum(-1)
    m = (e > e.mean(axis=(1, 2), keepdims=True)).astype(np.uint8)
    m[: n // 3] = rng.random((n // 3, SRC_H, SRC_W)) > 0.5
    m[0] = 0
    m[1] = 1
    m[2, ::2] = 1
    return m


def _packed_resize(mask, h, w):
    """What resize_mask_packed_kernel computes, instruction for instruction in uint32 arithmetic."""
    fx, fy, xd, yd = _geom(h, w)
    t_half = xd * yd // 2
    k = np.uint32(32767 - t_half)
    kk = np.uint32(k | (k << np.uint32(16)))
    x0, x1, xn, xden = oracle._linear_taps_exact(SRC_W, w)
    assert xden == 2 * w and np.all(xn % fx == 0)
    m = (np.asarray(mask) != 0).astype(np.uint32)
    rows = m[:, x0] * (xd - xn // fx).astype(np.uint32)[None, :] + m[:, x1] * (xn // fx).astype(np.uint32)[None, :]
    rows = np.concatenate([rows[:1], rows, rows[-1:]], 0)                 # one copy of the edge rows on either side
    wp = (w + 3) & ~3
    rows = np.pad(rows, ((0, 0), (0, wp - w)))
    out = np.zeros((h, w), np.uint8)
    for y in range(h):
        t = (2 * y + 1) * SRC_H - h
        lo = t // (2 * h)                                                  # floor, unclamped: -1 .. 35
        n = np.uint32((t - lo * 2 * h) // fy)
        a, b = rows[lo + 1], rows[lo + 2]
        pa = (a[0::2] | (a[1::2] << np.uint32(16))).astype(np.uint32)     # two pixels per register
        pb = (b[0::2] | (b[1::2] << np.uint32(16))).astype(np.uint32)
        acc = pa * np.uint32(yd - n) + kk
        acc = pb * n + acc                                                 # no carry between the halves, no wrap
        flags = np.stack([(acc >> np.uint32(15)) & np.uint32(1), (acc >> np.uint32(31)) & np.uint32(1)], 1).reshape(-1)
        out[y] = flags[:w]
    return out


@pytest.mark.parametrize('h,w', SIZES)
def test_reduced_taps_fit_sixteen_bits_and_repeat(h, w):
    fx, fy, xd, yd = _geom(h, w)
    assert (xd * yd) % 2 == 0 and xd * yd <= 65535                         # the static_asserts of PackedGeom
    y0, y1, yn, yden = oracle._linear_taps_exact(SRC_H, h)
    assert yden == 2 * h and np.all(yn % fy == 0)
    period_out, period_src = h // gcd(h, SRC_H), SRC_H // gcd(h, SRC_H)
    assert period_out == 56 and period_src == 9
    lo = np.array([((2 * y + 1) * SRC_H - h) // (2 * h) for y in range(h)])
    n = np.array([((2 * y + 1) * SRC_H - h) - lo[y] * 2 * h for y in range(h)]) // fy
    assert lo[0] == -1 and lo[-1] == SRC_H - 1
    assert np.array_equal(lo[period_out:], lo[:-period_out] + period_src) and np.array_equal(n[period_out:], n[:-period_out])
    inner = (lo >= 0) & (lo < SRC_H - 1)                                   # away from the clamped border rows the taps agree
    assert np.array_equal(lo[inner], y0[inner]) and np.array_equal(n[inner], yn[inner] // fy)


@pytest.mark.parametrize('h,w', SIZES)
def test_packed_resize_arithmetic_equals_oracle(h, w):
    for m in _masks(24, 3):
        assert np.array_equal(_packed_resize(m, h, w), oracle.resize_mask(m, h, w))


@pytest.mark.parametrize('h,w', SIZES)
def test_rectangle_count_form_of_consensus_iou_equals_oracle(h, w):
    """2 I = P(B0) + P(B1) + P(B2) - P(B0 n B1 n B2); 2 U = areas(B0, B1, B2) - area(B0 n B1 n B2) + 2 [P(all) - P(B0 u B1 u B2)]."""
    n = 48
    masks = _masks(n, 4)
    rng = np.random.default_rng(5)
    xmin, xmax, ymin, ymax = [a.copy() for a in synth.flickr_boxes(n, 6, h, w)]
    xmin[:8] = xmin[:8, :1]; xmax[:8] = xmax[:8, :1]; ymin[:8] = ymin[:8, :1]; ymax[:8] = ymax[:8, :1]      # identical triples
    for i in range(8, 20):                                                                                   # any order / range
        a, b = rng.integers(-20, w + 20, (3, 2)), rng.integers(-20, h + 20, (3, 2))
        xmin[i], xmax[i], ymin[i], ymax[i] = a[:, 0], a[:, 1], b[:, 0], b[:, 1]
    want_i, want_u, _, _ = oracle.flickr_sweep(masks, xmin, xmax, ymin, ymax, [0.5], h, w)

    def clip(i, c):
        xa, xb = max(min(xmin[i, c], xmax[i, c]), 0), min(max(xmin[i, c], xmax[i, c]), w - 1)
        ya, yb = max(min(ymin[i, c], ymax[i, c]), 0), min(max(ymin[i, c], ymax[i, c]), h - 1)
        return None if xmax[i, c] == 0 or xa > xb or ya > yb else (xa, xb, ya, yb)

    def meet(p, q):
        if p is None or q is None:
            return None
        r = (max(p[0], q[0]), min(p[1], q[1]), max(p[2], q[2]), min(p[3], q[3]))
        return None if r[0] > r[1] or r[2] > r[3] else r

    for i in range(n):
        pred = oracle.resize_mask(masks[i], h, w).astype(np.int64)
        count = lambda r: 0 if r is None else int(pred[r[2]:r[3] + 1, r[0]:r[1] + 1].sum())
        area = lambda r: 0 if r is None else (r[1] - r[0] + 1) * (r[3] - r[2] + 1)
        b = [clip(i, c) for c in range(3)]
        b01, b02, b12 = meet(b[0], b[1]), meet(b[0], b[2]), meet(b[1], b[2])
        b012 = meet(b01, b[2])
        i2 = sum(count(r) for r in b) - count(b012)
        covered = sum(count(r) for r in b) - count(b01) - count(b02) - count(b12) + count(b012)
        u2 = sum(area(r) for r in b) - area(b012) + 2 * (int(pred.sum()) - covered)
        assert (i2, u2) == (int(want_i[i]), int(want_u[i])), i
