"""Round-2 kernels through the C ABI, against the oracle, the golden vectors and each other.  Needs a B200.

* stage2_kernel / stage2_cluster_kernel (one thread per pixel; one CTA or one cluster of 8 CTAs per frame): the two
  forms must agree bit for bit with each other - energies, means, masks, in-place scaling - and with the oracle within
  the float64 tolerance; NaN propagation of the min / max; in-place aliasing on device tensors.
* aig_acivw_batch: the reference's evaluation step in one launch == energy + energy + iou_sweep, == the golden replay.
* heat_stream_kernel (bulk shared -> global copies): == the round-1 per-thread-store kernel bit for bit, fused
  aig_energy_heatmap == the two-kernel chain bit for bit.
* debug_jitter: 200 random (frames, seed) cases of the jittered persistent kernel == the sequential two-kernel path.
* one-rank NCCL communicator: aig_comm_init / aig_allreduce_counts run NCCL on a single-GPU box.
"""
import ctypes

import numpy as np
import pytest

import acoustic_image_generation_b200 as aig
from acoustic_image_generation_b200 import evaluate, synth
from oracle import acoustic_oracle as oracle

pytestmark = pytest.mark.gpu

ENERGY_RTOL = 1e-13
REF_THR = list(oracle.REFERENCE_THRESHOLDS)


@pytest.fixture(scope='module')
def path():
    p = aig.AcousticPath(0)
    yield p
    p.close()


@pytest.fixture(scope='module')
def path_cta():
    """A handle that never takes the cluster-per-frame form: one CTA per frame whatever the batch size."""
    p = aig.AcousticPath(0)
    p.set_option('small_batch_frames', 1)
    yield p
    p.close()


@pytest.fixture(scope='module')
def torch():
    import torch
    assert torch.cuda.is_available()
    return torch


def _images_with_hard_cases(n, seed):
    """Smooth / sigmoid images plus the pixels the fast float64 path hands to the plain path: huge values (|mel| > 700),
    an Inf, a NaN, a constant frame."""
    imgs = np.concatenate([synth.smooth_images(n // 2, seed), synth.sigmoid_images(n - n // 2, seed + 1)], 0)
    imgs[1, 3, 5, :] = 400.0                 # exp overflow on the fast path -> plain path
    imgs[2, 10, 7, 4] = np.inf
    imgs[3, 0, 0, 0] = -250.0
    imgs[4] = 0.25                           # constant frame
    return imgs


# ----------------------------------------------------------------------------------------------
# energy: the two kernel forms, the oracle, aliasing, NaN
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize('normalize_first', [False, True])
def test_energy_cluster_and_cta_forms_are_bit_identical(path, path_cta, normalize_first):
    imgs = _images_with_hard_cases(23, 40)
    a = path.energy(imgs, normalize_first=normalize_first, want_scaled=True, want_mean=True)
    b = path_cta.energy(imgs, normalize_first=normalize_first, want_scaled=True, want_mean=True)
    for name, x, y in zip(('energy', 'mask', 'scaled', 'mean'), a, b):
        assert np.array_equal(x, y, equal_nan=True), name
    want_e, want_m = oracle.energy_stage(imgs, normalize_first=normalize_first)
    ok = np.isfinite(want_e)
    assert np.array_equal(np.isfinite(a[0]), ok)
    assert np.abs(a[0][ok] - want_e[ok]).max() <= ENERGY_RTOL * np.abs(want_e[ok]).max()
    diff = int((a[1] != want_m).sum())
    print('masks differing from the oracle on the hard-case batch: %d' % diff)
    assert diff == 0


@pytest.mark.parametrize('normalize_first', [False, True])
def test_energy_wide_kernel_equals_one_cta_per_frame_kernel(path_cta, normalize_first):
    """aig_energy on large batches runs stage2_wide_kernel (eight frames per 512-thread CTA, eight interleaved copies of
    the exponential table); option energy_wide = 0 keeps stage2_kernel<1> (one 64-thread CTA per frame, one copy).  Same
    bits - energies, masks, in-place scaled values, means - for batch sizes around the group count and the grid size,
    hard cases (plain-path pixels, Inf, NaN, a constant frame) included; and the oracle's masks."""
    for n in (1, 7, 8, 9, 23, 300, 8 * 148 + 5):
        imgs = _images_with_hard_cases(max(n, 6), 50 + n)[:n] if n >= 6 else synth.smooth_images(n, 50 + n)
        try:
            path_cta.set_option('energy_wide', 1)
            a = path_cta.energy(imgs, normalize_first=normalize_first, want_scaled=True, want_mean=True)
            path_cta.set_option('energy_wide', 0)
            b = path_cta.energy(imgs, normalize_first=normalize_first, want_scaled=True, want_mean=True)
        finally:
            path_cta.set_option('energy_wide', 1)
        for name, x, y in zip(('energy', 'mask', 'scaled', 'mean'), a, b):
            assert np.array_equal(x, y, equal_nan=True), (n, name)
        if n <= 23:
            want_e, want_m = oracle.energy_stage(imgs, normalize_first=normalize_first)
            assert np.array_equal(a[1], want_m), n
    # the pair form (aig_acivw_batch: four pairs per CTA) against stage2_kernel<2>, 101 thresholds
    for n in (1, 3, 4, 5, 4 * 148 + 3):
        real = _images_with_hard_cases(max(n, 6), 70 + n)[:n]
        recon = synth.smooth_images(n, 71 + n)
        recon[0] = real[0]                                 # identical pair: IoU exactly 1 (or 0 / 0 for a NaN frame)
        out = []
        try:
            path_cta.set_option('acivw_wide_pairs', 0)     # the default takes the pair form from 8192 pairs up only
            for wide in (1, 0):
                path_cta.set_option('energy_wide', wide)
                out.append(path_cta.acivw_batch(real, recon, np.linspace(0, 1, 101), normalize_first=normalize_first,
                                                want_energy=True, want_masks=True))
        finally:
            path_cta.set_option('energy_wide', 1)
            path_cta.set_option('acivw_wide_pairs', 8192)
        (i1, u1, pos1, num1, e1, m1), (i0, u0, pos0, num0, e0, m0) = out
        assert np.array_equal(i1, i0) and np.array_equal(u1, u0) and np.array_equal(pos1, pos0) and num1 == num0 == n
        for x, y in zip(e1 + m1, e0 + m0):
            assert np.array_equal(x, y, equal_nan=True)


def test_energy_large_batch_equals_small_batch_form(path, path_cta):
    """Above the SM count the default handle runs one CTA per frame; below, a cluster per frame: same bits."""
    imgs = synth.smooth_images(200, 41)
    big = path.energy(imgs, normalize_first=True, want_mean=True)
    small = [path.energy(imgs[i:i + 50], normalize_first=True, want_mean=True) for i in range(0, 200, 50)]
    for k in range(3):
        assert np.array_equal(big[k], np.concatenate([s[k] for s in small], 0))


def test_find_logen_in_place_on_device_tensors_both_forms(path, path_cta, golden, torch):
    """scaled_out == images (find_logen scales its argument in place): ADVICE r1 flagged read-only loads on memory the
    kernel writes.  Both forms, one frame and a batch, against the reference-run golden."""
    g = golden('energy')
    for p in (path, path_cta):
        dev = torch.from_numpy(synth.sigmoid_images(2, 3)[0].copy()).cuda()
        en = p.find_logen(dev)
        assert np.array_equal(dev.cpu().numpy(), g['sigmoid0_after_call'])
        assert np.abs(en.cpu().numpy() - g['energy_sigmoid'][0]).max() <= ENERGY_RTOL * float(en.max())
        # a batch, in place through the raw ABI (images == scaled_out), hard cases included
        imgs = _images_with_hard_cases(12, 44)
        want = p.energy(imgs, want_scaled=True)
        t = torch.from_numpy(imgs.copy()).cuda()
        e = torch.empty((12, 36, 48), dtype=torch.float64, device='cuda')
        assert p._lib.aig_energy(p._h, t.data_ptr(), 12, 0, t.data_ptr(), e.data_ptr(), None, None) == 0
        assert np.array_equal(t.cpu().numpy(), want[2], equal_nan=True)
        assert np.array_equal(e.cpu().numpy(), want[0], equal_nan=True)


def test_nan_propagates_through_min_max_like_tf(path, path_cta):
    """tf.reduce_min / reduce_max propagate NaN (outdoor_data_mfcc.py:674,677): a frame holding one NaN normalises to
    all-NaN; its energies are NaN and its mask empty.  (fminf / fmaxf would have dropped the NaN.)"""
    imgs = synth.sigmoid_images(3, 45)
    imgs[1, 7, 9, 2] = np.nan
    for p in (path, path_cta):
        normed = p.normalize_images(imgs)
        assert np.isnan(normed[1]).all() and np.isfinite(normed[0]).all() and np.isfinite(normed[2]).all()
        assert np.array_equal(normed, oracle.normalize_acoustic_images(imgs), equal_nan=True)
        energy, mask = p.energy(imgs, normalize_first=True)
        assert np.isnan(energy[1]).all() and mask[1].sum() == 0
        assert np.isfinite(energy[0]).all() and np.isfinite(energy[2]).all()
    vec = synth.sigmoid_images(1, 46)[0, 0, :5, :].copy()           # five 12-vectors
    vec[2, 3] = np.nan
    got = path.normalize_mfcc(vec)
    assert np.isnan(got[2]).all() and np.isfinite(got[[0, 1, 3, 4]]).all()


def _normalisation_frames():
    """Frames built around the three forms of the per-frame division (FrameNormFast in csrc/energy_kernel.cuh): mode 2
    (no per-value test: |min| >= 2^-14 range), mode 1 (per-value test: minimum at or near zero) and mode 0 (IEEE division:
    tiny / huge range), each salted with values one or a few ulps above the minimum, exact zeros, the maximum itself and
    values straddling the 2^-14 bound."""
    rng = np.random.default_rng(77)
    frames = []

    def frame(lo, hi):
        f = rng.uniform(lo, hi, (36, 48, 12)).astype(np.float32)
        flat = f.reshape(-1)
        lo32, hi32 = np.float32(lo), np.float32(hi)
        flat[0], flat[1] = lo32, hi32
        v = lo32
        for i in range(2, 40):                      # the smallest non-zero differences the frame can hold
            v = np.nextafter(v, np.float32(np.inf), dtype=np.float32)
            flat[i] = v
        flat[40:60] = lo32 + (hi32 - lo32) * np.float32(2.0) ** -np.arange(20, 40).astype(np.float32)   # d / range down to 2^-39
        return f

    frames.append(frame(-37.5, 22.25))                                   # MFCC-like: mode 2
    frames.append(frame(3.0, 9.0))                                       # positive minimum: mode 2
    frames.append(frame(-2.0 ** -14, 1.0 - 2.0 ** -14))                  # |lo| == 2^-14 range: the bound itself
    frames.append(frame(-2.0 ** -14 * 0.999, 1.0))                       # just under it: mode 1
    frames.append(frame(0.0, 1.0))                                       # an already normalised frame: mode 1
    frames.append(frame(1e-30, 1.0))                                     # minimum a denormal-sized step from zero
    frames.append(frame(-1e-12, 3e-12))                                  # range below 2^-30: mode 1 / tiny quotients
    frames.append(frame(-3e-25, 1e-25))                                  # range below 2^-60: IEEE division
    frames.append(frame(-1e25, 2e25))                                    # range above 2^60: IEEE division
    frames.append(frame(5.0, 5.0 + 2.0 ** -20))                          # a handful of distinct values
    return np.stack(frames, 0)


def test_min_max_division_forms_are_the_ieee_quotient(path, path_cta, torch):
    """(x - min) / (max - min) of every value, bit for bit the float32 division the reference does (:672-679), through
    the kernels' reciprocal forms: frames for each FrameNormFast mode, read back as the scaled image - which is
    the quotient divided by the lifter and multiplied by mfnorm, so it is compared with the oracle's chain - and through
    the stand-alone normalise kernel (IEEE division per value)."""
    imgs = _normalisation_frames()
    want_norm = np.stack([oracle.normalize_acoustic_image(f) for f in imgs], 0)
    got_norm = path.normalize_images(imgs)
    assert np.array_equal(got_norm.view(np.uint32), want_norm.view(np.uint32))
    want_scaled = want_norm.copy()
    want_energy = np.stack([oracle.find_logen(f) for f in want_scaled], 0)          # scales want_scaled in place
    for p in (path, path_cta):
        energy, _, scaled = p.energy(torch.from_numpy(imgs.copy()).cuda(), normalize_first=True, want_scaled=True)
        got = scaled.cpu().numpy()
        differ = np.argwhere(got.view(np.uint32) != want_scaled.view(np.uint32))
        assert differ.size == 0, 'first differing values (frame, y, x, c): %s' % differ[:5].tolist()
        np.testing.assert_allclose(energy.cpu().numpy(), want_energy, rtol=ENERGY_RTOL, atol=0)


def test_normalise_bulk_copy_kernel(path, torch):
    """aig_normalize_images with the frame resident in shared memory between a bulk load and a bulk store
    (normalize_bulk_kernel: two buffers per CTA, several frames per CTA) == the per-thread kernel == the oracle, bit for
    bit: 700 frames (4-5 per CTA on 148 SMs) including the adversarial ones, a NaN frame and a constant frame; in place;
    device buffers off the 16-byte grid are refused, here and by aig_tile_mfcc."""
    rng = np.random.default_rng(12)
    imgs = np.concatenate([_normalisation_frames(), (rng.normal(-8, 12, (690, 36, 48, 12))).astype(np.float32)], 0)
    imgs[13, 5, 5, 5] = np.nan
    imgs[14] = 2.5
    want = oracle.normalize_acoustic_images(imgs)
    d = torch.from_numpy(imgs).cuda()
    got = path.normalize_images(d)

    def same(a, b):                                  # bit for bit, any NaN equal to any NaN (payloads are not specified)
        a, b = np.asarray(a), np.asarray(b)
        nan = np.isnan(a)
        return np.array_equal(nan, np.isnan(b)) and np.array_equal(a[~nan].view(np.uint32), b[~nan].view(np.uint32))

    got_host = got.cpu().numpy()
    assert np.isnan(got_host[13]).all() and np.isnan(got_host[14]).all()
    assert same(got_host, want)
    path.set_option('norm_bulk_copy', 0)
    try:
        old = path.normalize_images(d)
    finally:
        path.set_option('norm_bulk_copy', 1)
    assert same(old.cpu().numpy(), got_host)
    lib, h = path._lib, path._h
    assert lib.aig_normalize_images(h, d.data_ptr(), d.shape[0], d.data_ptr()) == 0          # in place
    torch.cuda.synchronize()
    assert same(d.cpu().numpy(), got_host)
    flat = torch.zeros(2 * 20736 + 8, device='cuda')
    out = torch.empty(2 * 20736 + 8, device='cuda')
    assert lib.aig_normalize_images(h, flat.data_ptr() + 4, 2, out.data_ptr()) == -1         # AIG_ERR_ARGUMENT, no fault
    assert b'aligned' in lib.aig_last_error(h)
    assert lib.aig_tile_mfcc(h, flat.data_ptr() + 4, 2, 0, out.data_ptr()) == -1
    assert b'aligned' in lib.aig_last_error(h)
    assert lib.aig_normalize_images(h, flat.data_ptr(), 2, out.data_ptr()) == 0              # the handle still works
    torch.cuda.synchronize()


def test_selftest_hoisted_reciprocal_division(path):
    """FrameNormFast::apply against __fdiv_rn on 2^32 (value, range) pairs: no difference anywhere."""
    bad, count, fast, _ = path.selftest(2)
    print('min-max division selftest: %d pairs, %d on the reciprocal path, %d mismatches' % (count, fast, bad))
    assert count == 1 << 32 and bad == 0 and fast > count // 3


# ----------------------------------------------------------------------------------------------
# aig_acivw_batch
# ----------------------------------------------------------------------------------------------
def test_acivw_batch_matches_reference_replay(path, path_cta, golden):
    g = golden('acivw_iou')
    n = int(g['num'])
    a, b = synth.smooth_images(n, 10), synth.smooth_images(n, 11)
    b[: n // 2] = a[: n // 2] * np.float32(0.9) + b[: n // 2] * np.float32(0.1)
    assert synth.digest(a) == str(g['digest_a']) and synth.digest(b) == str(g['digest_b'])
    keep_a, keep_b = a.copy(), b.copy()
    for p in (path, path_cta):
        inter, union, pos, num = p.acivw_batch(a, b)
        assert np.array_equal(inter, g['inter']) and np.array_equal(union, g['union'])
        assert np.array_equal(pos, g['pos11']) and num == n
        inter, union, pos, num = p.acivw_batch(a, b, np.linspace(0, 1, 101))
        assert np.array_equal(pos, g['pos101'])
    assert np.array_equal(a, keep_a) and np.array_equal(b, keep_b)      # the reference works on an np.stack copy


@pytest.mark.parametrize('n', [1, 7, 16, 150, 333])
@pytest.mark.parametrize('normalize_first', [False, True])
def test_acivw_batch_equals_separate_kernels(path, path_cta, torch, n, normalize_first):
    a = _images_with_hard_cases(max(n, 6), 50 + n)[:n] if n >= 6 else synth.smooth_images(n, 50)
    b = synth.smooth_images(n, 51 + n)
    if n > 3:
        b[3] = a[3]                                        # identical pair: IoU exactly 1
    e_a, m_a = path.energy(a, normalize_first=normalize_first)
    e_b, m_b = path.energy(b, normalize_first=normalize_first)
    want_i, want_u, want_pos, want_num = path.iou_sweep(m_a, m_b, REF_THR)
    for p in (path, path_cta):
        inter, union, pos, num, energies, masks = p.acivw_batch(a, b, REF_THR, normalize_first=normalize_first,
                                                                want_energy=True, want_masks=True)
        assert np.array_equal(inter, want_i) and np.array_equal(union, want_u)
        assert np.array_equal(pos, want_pos) and num == want_num == n
        assert np.array_equal(energies[0], e_a, equal_nan=True) and np.array_equal(energies[1], e_b, equal_nan=True)
        assert np.array_equal(masks[0], m_a) and np.array_equal(masks[1], m_b)
    # device tensors, accumulating device counters (two calls)
    thr = torch.tensor(REF_THR, dtype=torch.float64, device='cuda')
    counts = torch.zeros(12, dtype=torch.int64, device='cuda')
    for _ in range(2):
        path.acivw_batch(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), thr, pos=counts[:-1], num=counts[-1:],
                         normalize_first=normalize_first)
    assert np.array_equal(counts.cpu().numpy(), np.concatenate([2 * want_pos, [2 * n]]))


def test_acivw_evaluation_driver_uses_one_launch_per_batch(path, golden):
    g = golden('acivw_iou')
    n = int(g['num'])
    a, b = synth.smooth_images(n, 10), synth.smooth_images(n, 11)
    b[: n // 2] = a[: n // 2] * np.float32(0.9) + b[: n // 2] * np.float32(0.1)
    ev = evaluate.AcivwEvaluation(path)
    before = path.launch_count
    for lo in range(0, n, 16):                            # the reference's batch size
        ev.add_batch(a[lo:lo + 16], b[lo:lo + 16])
    assert path.launch_count - before == 3                # one kernel per add_batch
    res = ev.finish()
    assert np.array_equal(res['pos'], g['pos11']) and res['num'] == n
    assert abs(res['auc_exact'] - float(golden('auc')['acivw11'])) <= 1e-12
    assert abs(res['auc'] - float(golden('auc_files')['acivw11_auc_from_files'])) <= 1e-12


def test_consensus_iou_exact_tenth_is_not_counted(path, golden):
    """I / U exactly 1/10 and 3/10: the reference's ratio is float64 (golden made from its statements), so thresholds 0.1
    and 0.3 do not count those frames; the kernel's float64 ratio agrees."""
    g = golden('ciou_ratio')
    i2, u2, pos, num = path.ciou_sweep(g['masks'], g['xmin'], g['xmax'], g['ymin'], g['ymax'], REF_THR, out_hw=(36, 48))
    assert np.array_equal(i2, (2 * g['inter']).astype(np.int64)) and np.array_equal(u2, (2 * g['union']).astype(np.int64))
    assert np.array_equal(pos, g['pos11']) and num == 3


# ----------------------------------------------------------------------------------------------
# heat maps: bulk-copy kernel, fused energy + heat map
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize('shape', [(224, 298), (224, 224), (36, 48), (100, 78), (448, 596), (7, 12), (2, 2)])
def test_heat_stream_kernel_equals_per_thread_store_kernel(path, shape):
    """Same arithmetic (float32 lerp, min / max on the edge rows, (v - min) * inv), different way out of the SM: staged rows
    + cp.async.bulk instead of st.global.  Bit-identical, and within tolerance of the oracle."""
    rng = np.random.default_rng(shape[0] * 31 + shape[1])
    e = rng.random((301, 36, 48)) ** 3
    e[5] = 0.75                                           # constant map -> NaN
    got = path.heatmap(e, *shape)
    path.set_option('heat_bulk_store', 0)
    try:
        old = path.heatmap(e, *shape)
    finally:
        path.set_option('heat_bulk_store', 1)
    assert np.array_equal(got, old, equal_nan=True)
    assert np.isnan(got[5]).all()
    for i in (0, 150, 300):
        assert np.abs(got[i] - oracle.heatmap(e[i], *shape)).max() <= 2e-6
        assert got[i].min() == 0.0 and abs(float(got[i].max()) - 1.0) <= 1e-6


@pytest.mark.parametrize('shape', [(224, 298), (224, 224)])
@pytest.mark.parametrize('normalize_first', [False, True])
def test_fused_energy_heatmap_equals_the_two_kernel_chain(path, torch, shape, normalize_first):
    """n >= SM count takes the single-launch kernel (energy map handed over in shared memory); the chain of aig_energy
    and aig_heatmap is the reference for it: energies, masks and heat maps bit for bit."""
    imgs = _images_with_hard_cases(400, 60)
    before = path.launch_count
    energy, mask, heat = path.energy_heatmap(imgs, normalize_first, *shape)
    assert path.launch_count - before == 1
    e2, m2 = path.energy(imgs, normalize_first=normalize_first)
    h2 = path.heatmap(e2, *shape)
    assert np.array_equal(energy, e2, equal_nan=True) and np.array_equal(mask, m2)
    assert np.array_equal(heat, h2, equal_nan=True)
    _, _, only_heat = path.energy_heatmap(torch.from_numpy(imgs).cuda(), normalize_first, *shape, want_energy=False, want_mask=False)
    assert np.array_equal(only_heat.cpu().numpy(), heat, equal_nan=True)
    i = 7
    base = imgs[i:i + 1]
    want = oracle.heatmap(oracle.energy_stage(base, normalize_first=normalize_first)[0][0], *shape)
    assert np.abs(heat[i] - want).max() <= 1e-4


@pytest.mark.parametrize('shape', [(224, 298), (224, 224), (100, 78), (64, 60), (448, 596)])
@pytest.mark.parametrize('n_frames', [149, 297, 1000])
def test_warp_specialised_energy_heatmap_equals_the_sequential_kernel(path, shape, n_frames):
    """energy_heat_ws_kernel (float64 warps and heat-map warps of one CTA hand the map over through mbarriers; default) against
    heat_stream_kernel<true> (option energy_heat_ws = 0): energies, masks and heat maps bit for bit, for frame counts that
    leave CTAs with 0, 1 and several frames and an odd one out, a generic-size build and a size too large for two CTAs
    per SM (falls back).  The hand-over is also a race test: any frame read too early or overwritten too late shows."""
    imgs = _images_with_hard_cases(n_frames, 40)
    for normalize_first in (False, True):
        before = path.launch_count
        got = path.energy_heatmap(imgs, normalize_first, *shape)
        assert path.launch_count - before == (1 if shape[0] <= 224 else 2)     # 448 x 596 rows do not fit: two kernels
        path.set_option('energy_heat_ws', 0)
        try:
            want = path.energy_heatmap(imgs, normalize_first, *shape)
        finally:
            path.set_option('energy_heat_ws', 1)
        for g, w in zip(got, want):
            assert np.array_equal(g, w, equal_nan=True)


@pytest.mark.parametrize('shape', [(224, 298), (224, 224), (100, 78), (37, 49)])
def test_persistent_kernel_with_heat_map_phase(path, torch, shape):
    """aig_mfcc_energy_heatmap (opt-in): the energy warps of the persistent kernel also up-sample and normalise.  MFCC,
    energy and mask must equal aig_mfcc_energy's and the heat map aig_heatmap's, bit for bit; (37, 49) cannot stream
    (odd width) and takes the separate kernels, 3 frames the small-batch path."""
    pool = torch.from_numpy(synth.power_frames(12, 73, 'chi2')).cuda()
    for n in (3, 310):
        power = pool[torch.arange(n, device='cuda') % 12].contiguous()
        mfcc, energy, mask = path.mfcc_energy(power, flip=True, normalize_first=True)
        want_heat = path.heatmap(energy, *shape)
        before = path.launch_count
        got = path.mfcc_energy_heatmap(power, flip=True, normalize_first=True, out_h=shape[0], out_w=shape[1])
        launches = path.launch_count - before
        assert torch.equal(got[0], mfcc) and torch.equal(got[1], energy) and torch.equal(got[2], mask)
        assert torch.equal(got[3], want_heat)
        if n == 310 and shape[1] % 2 == 0:
            assert launches == 1
    host = path.mfcc_energy_heatmap(pool[:2].cpu().numpy(), flip=True, normalize_first=True, out_h=shape[0], out_w=shape[1])
    dev = path.mfcc_energy_heatmap(pool[:2].contiguous(), flip=True, normalize_first=True, out_h=shape[0], out_w=shape[1])
    assert np.array_equal(host[3], dev[3].cpu().numpy()) and np.array_equal(host[0], dev[0].cpu().numpy())


def test_heatmap_exact_mode_still_bit_exact_through_the_combined_call(path):
    imgs = synth.smooth_images(160, 61)
    path.set_option('heatmap_exact', 1)
    try:
        energy, _, heat = path.energy_heatmap(imgs)
    finally:
        path.set_option('heatmap_exact', 0)
    for i in (0, 80, 159):
        assert np.array_equal(heat[i], oracle.heatmap(energy[i]).astype(np.float32))


# ----------------------------------------------------------------------------------------------
# small batches of the chained call, race stress
# ----------------------------------------------------------------------------------------------
def test_small_batch_mfcc_energy_equals_persistent_kernel(path, path_cta):
    """Below the SM count aig_mfcc_energy runs the tiled MFCC kernel + the cluster energy kernel; path_cta
    (small_batch_frames = 1) keeps the one-CTA-per-frame persistent kernel.  Same bits."""
    power = synth.power_frames(9, 70, 'chi2')
    a = path.mfcc_energy(power, flip=True, normalize_first=True, want_mean=True)
    b = path_cta.mfcc_energy(power, flip=True, normalize_first=True, want_mean=True)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_race_stress_jittered_persistent_kernel(torch):
    """compute-sanitizer's racecheck is closed on this pool.  Instead the persistent kernel is built a second time with
    pseudo-random clock64() spins before every mbarrier wait / arrive of its three roles (TMA producer, MFCC consumers,
    energy warps; option debug_jitter = seed): 200 random (frame count, seed, flip, normalise) cases must reproduce the
    sequential two-kernel path (chain_mode 0) bit for bit.  A hand-off that only works by timing luck would not."""
    rng = np.random.default_rng(2024)
    pool = torch.from_numpy(synth.power_frames(24, 71, 'chi2')).cuda()
    extra = torch.from_numpy(synth.power_frames(8, 72, 'lognormal')).cuda()
    pool = torch.cat([pool, extra], 0)
    ref_path = aig.AcousticPath(0)
    ref_path.set_option('chain_mode', 0)
    ref_path.set_option('small_batch_frames', 1)
    jit_path = aig.AcousticPath(0)
    failures = []
    try:
        for case in range(200):
            n = int(rng.integers(1, 420))
            seed = int(rng.integers(1, 2 ** 31 - 1))
            flip, norm = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
            idx = torch.from_numpy(rng.integers(0, len(pool), n)).cuda()
            power = pool[idx].contiguous()
            want = ref_path.mfcc_energy(power, flip=flip, normalize_first=norm, want_mean=True)
            jit_path.set_option('debug_jitter', seed)
            got = jit_path.mfcc_energy(power, flip=flip, normalize_first=norm, want_mean=True)
            for name, x, y in zip(('mfcc', 'energy', 'mask', 'mean'), got, want):
                if not torch.equal(x, y):
                    failures.append((case, n, seed, name))
    finally:
        ref_path.close()
        jit_path.close()
    assert not failures, failures[:10]


def test_race_stress_jittered_energy_heat_ws_kernel(torch):
    """The warp-specialised aig_energy_heatmap kernel hands each energy map from its float64 warps to its heat-map warps
    through one shared-memory slot and a full / empty mbarrier pair (and the slot doubles as the heat-map warps' t[]).  Its
    jittered build spins both sides for pseudo-random times around every wait / arrive; 120 random (frame count, seed, size,
    normalise) cases must reproduce the sequential kernel (energy_heat_ws = 0) bit for bit: energies, masks, heat maps."""
    rng = np.random.default_rng(77)
    pool = torch.from_numpy(_images_with_hard_cases(64, 90)).cuda()
    ref_path = aig.AcousticPath(0)
    ref_path.set_option('energy_heat_ws', 0)
    jit_path = aig.AcousticPath(0)
    failures = []
    try:
        for case in range(120):
            n = int(rng.integers(148, 700))
            seed = int(rng.integers(1, 2 ** 31 - 1))
            shape = [(224, 298), (224, 224), (96, 130), (50, 64)][int(rng.integers(0, 4))]
            norm = bool(rng.integers(0, 2))
            imgs = pool[torch.from_numpy(rng.integers(0, len(pool), n)).cuda()].contiguous()
            want = ref_path.energy_heatmap(imgs, norm, *shape)
            jit_path.set_option('debug_jitter', seed)
            before = jit_path.launch_count
            got = jit_path.energy_heatmap(imgs, norm, *shape)
            assert jit_path.launch_count - before == 1
            for name, x, y in zip(('energy', 'mask', 'heat'), got, want):
                same = torch.equal(x.view(torch.int64) if x.dtype == torch.float64 else x.view(torch.int32) if x.dtype == torch.float32 else x,
                                   y.view(torch.int64) if y.dtype == torch.float64 else y.view(torch.int32) if y.dtype == torch.float32 else y)
                if not same:
                    failures.append((case, n, seed, shape, name))
    finally:
        ref_path.close()
        jit_path.close()
    assert not failures, failures[:10]


# ----------------------------------------------------------------------------------------------
# NCCL on one GPU
# ----------------------------------------------------------------------------------------------
def test_one_rank_nccl_communicator_runs_the_native_allreduce(path, torch):
    """aig_comm_unique_id -> aig_comm_init(rank 0 of 1) -> aig_allreduce_counts: a legal one-rank NCCL communicator, so the
    library's own collective path (dlopen'd NCCL, the handle's stream, device and host buffers) is exercised on a
    single-GPU box; with several ranks only the sum changes (tests/test_gpu_multirank.py)."""
    p = aig.AcousticPath(0)
    try:
        ident = aig.AcousticPath.comm_unique_id()
        assert len(ident) == 128 and any(ident)
        p.init_comm_from_id(ident, 0, 1)
        assert p.has_comm
        dev = torch.arange(12, dtype=torch.int64, device='cuda') * 3
        before = dev.clone()
        p.allreduce_counts(dev)
        p.synchronize()
        assert torch.equal(dev, before)
        host = np.arange(12, dtype=np.int64) + 5
        p.allreduce_counts(host)
        assert np.array_equal(host, np.arange(12) + 5)
        # an evaluation that finishes through the handle's communicator
        ev = evaluate.AcivwEvaluation(p)
        ev.add_batch(synth.smooth_images(4, 80), synth.smooth_images(4, 81))
        res = ev.finish()
        assert res['num'] == 4
        assert p._lib.aig_comm_destroy(p._h) == 0
        bad = (ctypes.c_uint8 * 128)()
        assert p._lib.aig_comm_init(p._h, bad, 1, 1) == -1              # rank out of range
    finally:
        p.close()


# ----------------------------------------------------------------------------------------------
# packed mask kernels (mask_packed_kernel.cuh): the reference's two output sizes
# ----------------------------------------------------------------------------------------------
def _random_masks(n, seed):
    """Blobs (mean-threshold masks of smooth maps), salt-and-pepper, stripes, empty and full frames."""
    rng = np.random.default_rng(seed)
    imgs = synth.smooth_images(n, seed)
    e = imgs.sum(-1)
    m = (e > e.mean(axis=(1, 2), keepdims=True)).astype(np.uint8)
    m[n // 4:n // 2] = rng.random((n // 2 - n // 4, 36, 48)) > 0.5
    m[0] = 0
    m[1] = 1
    m[2, ::2] = 1
    m[2, 1::2] = 0
    m[3, :, ::2] = 1
    m[3, :, 1::2] = 0
    m[4] = 7                                               # any non-zero byte counts as set
    return m


def _adversarial_boxes(n, seed, h, w):
    xmin, xmax, ymin, ymax = [a.copy() for a in synth.flickr_boxes(n, seed, h, w)]
    rng = np.random.default_rng(seed + 1)
    k = n // 8
    # identical triples, nested, disjoint, swapped corners, out of range, 8-row-chunk boundaries, one-pixel boxes
    xmin[:k] = xmin[:k, :1]; xmax[:k] = xmax[:k, :1]; ymin[:k] = ymin[:k, :1]; ymax[:k] = ymax[:k, :1]
    for i in range(k, 2 * k):
        xmin[i] = [10, 20, 30]; xmax[i] = [w - 10, w - 20, w - 30]; ymin[i] = [8, 16, 23]; ymax[i] = [h - 9, h - 16, h - 24]
    for i in range(2 * k, 3 * k):
        a = rng.integers(-20, w + 20, (3, 2)); b = rng.integers(-20, h + 20, (3, 2))
        xmin[i], xmax[i], ymin[i], ymax[i] = a[:, 0], a[:, 1], b[:, 0], b[:, 1]          # any order, any range, xmax may be 0
    for i in range(3 * k, 4 * k):
        x, y = rng.integers(0, w, 3), rng.integers(0, h, 3)
        xmin[i], xmax[i], ymin[i], ymax[i] = x, np.maximum(x, 1), y, y                  # single pixels / columns
    return xmin, xmax, ymin, ymax


@pytest.mark.parametrize('shape', [(224, 298), (224, 224)])
def test_packed_resize_mask_equals_generic_kernel_and_oracle(path, torch, shape):
    masks = _random_masks(333, 11)
    before = path.launch_count
    got = path.resize_mask(masks, *shape)
    assert path.launch_count - before == 1
    path.set_option('mask_packed', 0)
    try:
        want = path.resize_mask(masks, *shape)
    finally:
        path.set_option('mask_packed', 1)
    assert got.dtype == np.uint8 and np.array_equal(got, want)
    for i in (0, 1, 2, 3, 4, 100, 332):
        assert np.array_equal(got[i], oracle.resize_mask(masks[i], *shape))
    # a device buffer that is not 16-byte aligned takes the generic kernel: same bytes
    dev = torch.from_numpy(masks).cuda()
    out = torch.zeros(333 * shape[0] * shape[1] + 4, dtype=torch.uint8, device='cuda')
    view = out[4:].view(333, *shape)
    path._check(path._lib.aig_resize_mask(path._h, dev.data_ptr(), 333, shape[0], shape[1], view.data_ptr()))
    assert np.array_equal(view.cpu().numpy(), got)


@pytest.mark.parametrize('shape', [(224, 298), (224, 224)])
def test_packed_ciou_equals_generic_kernel_and_oracle(path, shape):
    n = 640
    masks = _random_masks(n, 12)
    boxes = _adversarial_boxes(n, 13, *shape)
    thr = np.linspace(0, 1, 101)
    got = path.ciou_sweep(masks, *boxes, thr, out_hw=shape)
    path.set_option('mask_packed', 0)
    try:
        want = path.ciou_sweep(masks, *boxes, thr, out_hw=shape)
    finally:
        path.set_option('mask_packed', 1)
    for g, w in zip(got[:3], want[:3]):
        assert np.array_equal(g, w)
    assert got[3] == want[3] == n
    sel = np.r_[0:8, n // 8 - 2:n // 8 + 2, n // 4:n // 4 + 4, 3 * n // 8 - 2:3 * n // 8 + 2, n - 4:n]
    wi, wu, _, _ = oracle.flickr_sweep(masks[sel], *[b[sel] for b in boxes], thr, *shape)
    assert np.array_equal(got[0][sel], wi) and np.array_equal(got[1][sel], wu)
