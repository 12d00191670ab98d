"""The oracle restatement held to golden vectors produced by the reference's own source
(oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest

from acoustic_image_generation_b200 import synth
from oracle import acoustic_oracle as oracle

REF_THR = list(oracle.REFERENCE_THRESHOLDS)


def test_filterbank_bit_exact(golden):
    g = golden('filterbank')
    bank = oracle.createfilters(512, 24, 0, 6400, 12800)
    assert bank.dtype == np.float64 and bank.shape == (512, 24)
    assert np.array_equal(bank, g['filter_mat'])
    assert np.array_equal(oracle.createfilters(256, 20, 300, 4000, 8000), g['filter_mat_256_20'])
    # known answers recorded in SURVEY.md section 4
    assert synth.hashlib.sha256(bank.tobytes()).hexdigest().startswith('0d1ef0b07bb41770')
    assert int((bank != 0).sum()) == 942
    assert int((bank != 0).sum(1).max()) == 2
    assert not bank[0].any() and not bank[511].any()
    edges = oracle.mel_band_edges(512, 24, 0, 6400, 12800)
    assert edges.tolist() == [0, 5, 11, 17, 25, 32, 41, 51, 61, 72, 85, 99, 114, 130, 148, 168, 190,
                              214, 240, 269, 300, 335, 373, 415, 460, 511]


def test_constants_bit_exact(golden):
    g = golden('filterbank')
    dct_base, lifter, mfnorm = oracle.mfcc_constants()
    assert np.array_equal(dct_base, g['dct_base'])
    assert np.array_equal(lifter, g['lifter'])
    assert mfnorm == float(g['mfnorm'])


@pytest.mark.parametrize('kind,n', [('chi2', 2), ('lognormal', 1), ('floor', 1)])
def test_get_feats_matches_reference(golden, kind, n):
    g = golden('mfcc')
    power = synth.power_frames(n, int(g['seed_' + kind]), kind)
    assert synth.digest(power) == str(g['digest_' + kind]), 'synthetic generator drifted'
    got = oracle.mfcc_image(power)
    assert got.dtype == np.float32 and got.shape == (n, 36, 48, 12)
    # same NumPy calls on the same arrays: bit-exact
    assert np.array_equal(got, g['mfcc_' + kind])
    if kind == 'floor':
        assert np.all(np.abs(got) < 2e-5)   # DCT rows annihilate a constant log-floor vector up to f64 rounding


def test_get_feats_float64_rows(golden):
    g = golden('mfcc')
    bank, dct_base, lifter, mfnorm = oracle.reference_tables()
    rows = synth.power_frames(2, 0, 'chi2').reshape(-1, 512)[:64]
    got = oracle.get_feats(512, rows, 12, dct_base, mfnorm, lifter, bank)
    assert got.dtype == np.float64
    assert np.array_equal(got, g['mfcc_chi2_f64_rows'])


def test_get_feats_nan_inf_fixups():
    bank, dct_base, lifter, mfnorm = oracle.reference_tables()
    rows = np.ones((3, 512), np.float32)
    rows[0, 100] = np.nan
    rows[1, 200] = np.inf
    got = oracle.get_feats(512, rows, 12, dct_base, mfnorm, lifter, bank)
    assert np.isfinite(got).all()
    assert np.array_equal(got[0], np.zeros(12)) and np.array_equal(got[1], np.zeros(12))
    assert np.abs(got[2]).max() > 0


def test_flip180():
    x = np.arange(2 * 36 * 48 * 12, dtype=np.float32).reshape(2, 36, 48, 12)
    y = oracle.flip180(x)
    assert np.array_equal(y[1, 0, 0], x[1, 35, 47]) and np.array_equal(y[0, 35, 47], x[0, 0, 0])
    assert np.array_equal(oracle.flip180(y), x)
    flat = x.reshape(2, 1728, 12)
    assert np.array_equal(y.reshape(2, 1728, 12), flat[:, ::-1, :])


def test_minmax_and_energy_match_reference(golden):
    g = golden('energy')
    mf = golden('mfcc')['mfcc_chi2']
    normed = oracle.normalize_acoustic_images(mf)
    assert normed.dtype == np.float32
    assert np.array_equal(normed, g['normed_input'])
    assert normed.min() == 0.0 and normed.max() == 1.0
    for name, imgs in (('normed', normed), ('sigmoid', synth.sigmoid_images(2, 3)),
                       ('smooth', synth.smooth_images(2, 4))):
        if name != 'normed':
            assert synth.digest(imgs) == str(g['digest_' + name]), 'synthetic generator drifted'
        energy, mask = oracle.energy_stage(imgs, normalize_first=False)
        assert energy.dtype == np.float64
        assert np.array_equal(energy, g['energy_' + name])
        assert np.array_equal(mask, g['mask_' + name])
        assert np.array_equal(np.array([np.mean(e) for e in energy]), g['mean_' + name])


def test_find_logen_scales_argument_in_place(golden):
    g = golden('energy')
    frame = synth.sigmoid_images(2, 3)[0].copy()
    before = frame.copy()
    oracle.find_logen(frame)
    assert np.array_equal(frame, g['sigmoid0_after_call'])
    assert not np.array_equal(frame, before)
    keep = before.copy()
    oracle.find_logen(keep, inplace=False)
    assert np.array_equal(keep, before)


def test_constant_frame_normalises_to_nan():
    out = oracle.normalize_acoustic_image(np.full((36, 48, 12), 3.0, np.float32))
    assert np.isnan(out).all()


def test_resize_matches_cv2_golden(golden):
    g = golden('heatmap')
    e0 = golden('energy')['energy_smooth'][0]
    for shape, key in (((224, 298), 'up_224_298'), ((224, 224), 'up_224_224')):
        up = oracle.resize_bilinear(e0, *shape)
        scale = np.abs(g[key]).max()
        assert np.abs(up - g[key]).max() <= 1e-12 * scale
        hm = oracle.heatmap(e0, *shape)
        assert hm.dtype == np.float32 and hm.min() == 0.0 and hm.max() == 1.0


def test_mask_resize_bit_exact(golden):
    g = golden('heatmap')
    e = golden('energy')
    for key, masks, shape in (('mask_up_224_298', e['mask_smooth'], (224, 298)),
                              ('mask_up_224_224', e['mask_smooth'], (224, 224)),
                              ('mask_sigmoid_up_224_298', e['mask_sigmoid'], (224, 298))):
        want = np.unpackbits(g[key], axis=-1)[..., :shape[1]]
        got = np.stack([oracle.resize_mask(m, *shape) for m in masks], 0)
        assert np.array_equal(got, want)


def test_mask_resize_half_weight_columns():
    # 48 -> 298: output columns 74 and 223 sit exactly between two source columns
    m = np.zeros((36, 48), np.uint8)
    m[:, 12] = 1           # column 74 interpolates source columns 11 and 12 with weight 0.5
    up = oracle.resize_mask(m, 224, 298)
    assert up[:, 74].sum() == 0          # 0.5 is not > 0.5
    assert up[:, 75].sum() == 224


def test_acivw_iou_matches_reference(golden):
    g = golden('acivw_iou')
    n = int(g['num'])
    a = synth.smooth_images(n, 10)
    b = synth.smooth_images(n, 11)
    b[: n // 2] = a[: n // 2] * np.float32(0.9) + b[: n // 2] * np.float32(0.1)
    assert synth.digest(a) == str(g['digest_a']) and synth.digest(b) == str(g['digest_b'])
    ea, _ = oracle.energy_stage(a, normalize_first=False)
    eb, _ = oracle.energy_stage(b, normalize_first=False)
    for thr, key in ((REF_THR, 'pos11'), (np.linspace(0, 1, 101), 'pos101')):
        inter, union, pos, num = oracle.acivw_sweep(ea, eb, thr)
        assert np.array_equal(inter, g['inter']) and np.array_equal(union, g['union'])
        assert np.array_equal(pos, g[key]) and num == n
    scores = np.array([oracle.iou_pair(oracle.mean_mask(x), oracle.mean_mask(y))[2] for x, y in zip(ea, eb)])
    assert np.array_equal(scores, g['iou'])


def test_flickr_ciou_matches_reference(golden):
    g = golden('flickr_ciou')
    n = int(g['num'])
    pred = synth.smooth_images(n, 20)
    xmin, xmax, ymin, ymax = synth.flickr_boxes(n, 21)
    assert synth.digest(pred) == str(g['digest_pred'])
    assert synth.digest(np.stack([xmin, xmax, ymin, ymax])) == str(g['digest_boxes'])
    _, masks = oracle.energy_stage(pred, normalize_first=False)
    gt0 = oracle.boxes_to_consensus(xmin[0], xmax[0], ymin[0], ymax[0])
    assert np.array_equal(gt0, g['gt0'])
    assert np.array_equal(oracle.resize_mask(masks[0]), g['pred0'])
    for thr, key in ((REF_THR, 'pos11'), (np.linspace(0, 1, 101), 'pos101')):
        i2, u2, pos, num = oracle.flickr_sweep(masks, xmin, xmax, ymin, ymax, thr)
        assert np.array_equal(i2 / 2.0, g['inter']) and np.array_equal(u2 / 2.0, g['union'])
        assert np.array_equal(pos, g[key]) and num == n
    scores = []
    for h in range(n):
        gt = oracle.boxes_to_consensus(xmin[h], xmax[h], ymin[h], ymax[h])
        scores.append(oracle.consensus_iou(gt, oracle.resize_mask(masks[h]))[2])
    assert np.array_equal(np.array(scores), g['iou'])


def test_boxes_clip_and_absent():
    gt = oracle.boxes_to_consensus([0, 290, 0], [10, 298, 0], [0, 200, 0], [5, 224, 0])
    assert gt[0:6, 0:11].min() == 0.5 and gt[6, 0] == 0 and gt[0, 11] == 0
    assert gt[200:224, 290:298].min() == 0.5      # clipped at the right and bottom border
    assert gt.sum() == 0.5 * (6 * 11 + 24 * 8)
    three = oracle.boxes_to_consensus([0, 0, 0], [9, 9, 9], [0, 0, 0], [9, 9, 9])
    assert three.max() == 1.0 and three.sum() == 100.0   # 1.5 capped at 1


def test_auc_matches_sklearn(golden):
    g = golden('auc')
    a = golden('acivw_iou'); f = golden('flickr_ciou')
    assert oracle.auc(REF_THR, oracle.success_rates(a['pos11'], a['num'])) == float(g['acivw11'])
    assert oracle.auc(REF_THR, oracle.success_rates(f['pos11'], f['num'])) == float(g['flickr11'])
    got = oracle.auc(np.linspace(0, 1, 101), oracle.success_rates(f['pos101'], f['num']))
    assert abs(got - float(g['flickr101'])) <= 1e-15
    # decreasing x equals the ascending trapezoid
    v = np.array([1.0, 0.9, 0.7, 0.4, 0.1])
    t = np.linspace(0, 1, 5)
    assert abs(oracle.auc(t, v) - np.sum((t[1:] - t[:-1]) * (v[1:] + v[:-1]) / 2)) < 1e-15


def test_auc_the_way_the_reference_pipeline_computes_it(golden):
    """areaundercurve.py integrates the rates it parses back from the 'iou {:6f}' files (six decimals), not pos / num:
    golden = those files written and parsed with the reference's own statements, then sklearn's auc.  The host-side
    product function (aig.auc, plain C) and the oracle must both reproduce it; the unrounded AUC differs where the
    rates are not six-decimal numbers (sevenths, thirds)."""
    import acoustic_image_generation_b200 as aig
    g = golden('auc_files')
    for name in ('acivw11', 'flickr11', 'sevenths', 'thirds'):
        pos, num = g[name + '_pos'], int(g[name + '_num'])
        want = float(g[name + '_auc_from_files'])
        rates = oracle.success_rates(pos, num)
        assert abs(oracle.auc(oracle.REFERENCE_THRESHOLDS, oracle.rates_as_written(rates)) - want) <= 1e-12
        assert np.array_equal(aig.rates_as_written(rates), oracle.rates_as_written(rates))
        got = aig.auc(oracle.REFERENCE_THRESHOLDS, aig.rates_as_written(aig.success_rates(pos, num)))
        assert abs(got - want) <= 1e-12
        assert 'area {:6f}'.format(got) == str(g[name + '_area_text'])
    exact = oracle.auc(oracle.REFERENCE_THRESHOLDS, oracle.success_rates(g['sevenths_pos'], 7))
    assert abs(exact - float(g['sevenths_auc_from_files'])) > 1e-8          # the rounding is visible


def test_consensus_iou_ratio_is_float64(golden):
    """ADVICE (round 1) suspected the reference divides two float32 sums.  It does not: unionbig = union + (mtot - box)
    is float64 because box is int64 (showimages_bb.py:312-316), so iou_score = float32 / float64 -> float64, and a frame
    with I / U exactly 1/10 is NOT counted at threshold 0.1 (0.1 > 0.1 is false; in float32 it would be).  Golden =
    the reference's statements replayed on crafted masks (oracle/make_golden.py: replay_flickr_mask)."""
    g = golden('ciou_ratio')
    assert str(g['ratio_dtype']) == 'float64'
    assert g['iou'][0] == 0.1 and g['iou'][1] == 0.3
    assert not g['counted'][0][1] and not g['counted'][1][3]
    scores = []
    for f in range(3):
        gt = oracle.boxes_to_consensus(g['xmin'][f], g['xmax'][f], g['ymin'][f], g['ymax'][f], 36, 48)
        i2, u2, score = oracle.consensus_iou(gt, oracle.resize_mask(g['masks'][f], 36, 48))
        assert i2 == 2 * g['inter'][f] and u2 == 2 * g['union'][f] and score == g['iou'][f]
        scores.append(score)
    pos, num = oracle.success_counts(scores, oracle.REFERENCE_THRESHOLDS)
    assert np.array_equal(pos, g['pos11']) and num == 3


def test_success_counts_nan_and_strictness():
    pos, num = oracle.success_counts([np.nan, 0.5, 1.0, 0.0], [0.0, 0.5, 1.0])
    assert pos.tolist() == [2, 1, 0] and num == 4


# ---- "next" rows N1 / N2 ---------------------------------------------------------------------------
def test_tukey_window_matches_scipy(golden):
    from acoustic_image_generation_b200 import tables
    g = golden('audio_front')
    assert np.array_equal(oracle.tukey_window(), g['tukey'])
    assert np.array_equal(tables.tukey_window(), g['tukey'])


def test_audio_front_half_matches_reference(golden):
    g = golden('audio_front')
    audio_i = synth.audio_rows(24, 90, np.int32)
    audio_f = synth.audio_rows(8, 91, np.float32, amplitude=1.0)
    assert synth.digest(audio_i) == str(g['digest_int']) and synth.digest(audio_f) == str(g['digest_float'])
    assert np.array_equal(oracle.build_spectrograms(audio_i), g['mfcc_int'])
    assert np.array_equal(oracle.build_spectrograms(audio_f), g['mfcc_float'])
    assert np.array_equal(oracle.power_spectrum(audio_i, oracle.tukey_window())[:4], g['power_int_f64'])
    assert np.array_equal(oracle.butter_lowpass_filter(audio_i), g['lowpass_int'])
    assert np.array_equal(oracle.butter_lowpass_filter(audio_f), g['lowpass_float'])
    assert np.array_equal(oracle.build_spectrograms(g['lowpass_int']), g['mfcc_lowpassed_int'])


def test_normalize_and_tile_mfcc():
    rng = np.random.default_rng(0)
    m = rng.standard_normal((5, 12)).astype(np.float32) * 10
    n = oracle.normalize_mfcc(m)
    assert n.dtype == np.float32 and np.all(n.min(1) == 0) and np.all(n.max(1) == 1)
    t = oracle.tile_mfcc(m)
    assert t.shape == (5, 36, 48, 12) and np.array_equal(t[3, 17, 29], m[3]) and np.array_equal(t[:, 0, 0], m)
    assert np.isnan(oracle.normalize_mfcc(np.full((1, 12), 2.0, np.float32))).all()


def test_triplet_slices_and_losses():
    a, b = synth.sigmoid_images(3, 1), synth.sigmoid_images(3, 2)
    t = oracle.split_triplets(a)
    assert t.shape == (4, 3, 36, 48, 3) and np.array_equal(t[2], a[..., 6:9]) and t[1].flags.c_contiguous
    mse = oracle.triplet_mse(a, b)
    d = a.astype(np.float64) - b
    assert abs(mse[0] - np.mean(d * d)) <= 1e-8 and abs(mse[4] - np.mean(d[..., 9:] ** 2)) <= 1e-8
    assert abs(mse[0] - mse[1:].mean()) <= 1e-15


def test_overlay_gray_matches_cv2_and_jet_table_shape():
    import cv2
    from acoustic_image_generation_b200 import tables
    rng = np.random.default_rng(0)
    frame = rng.integers(0, 256, (224, 298, 3), dtype=np.uint8)
    f = frame.astype(np.int64)
    gray = (f[..., 0] * 3735 + f[..., 1] * 19235 + f[..., 2] * 9798 + 16384) >> 15
    assert np.array_equal(gray, cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY))
    lut = tables.jet_lut()
    assert lut.shape == (256, 3) and lut.dtype == np.uint8
    assert lut[0].tolist() == [0, 0, 127] and lut[255].tolist() == [127, 0, 0]      # dark blue -> dark red
    assert lut[128].tolist() == [124, 255, 121]                                   # green plateau in the middle
    out = oracle.overlay(np.zeros((4, 4), np.float32), None, lut)
    assert out.shape == (4, 4, 3) and np.array_equal(out[0, 0], lut[0])


def test_overlay_gray_level_is_an_integer_quotient():
    """The overlay kernels take the gray level of a luma as min(mulhi(256 (y - lo), M), 255) with one reciprocal M per
    frame (frontend_kernel.cuh: GrayLevels) instead of the oracle's float32 expression
    min(int(fl(fl((y - lo) / (hi - lo)) * 256)), 255): equal for every lo <= y <= hi in 0..255."""
    lo, hi, y = np.meshgrid(np.arange(256), np.arange(256), np.arange(256), indexing='ij', sparse=True)
    ok = (lo <= y) & (y <= hi) & (lo < hi)
    a = np.broadcast_to((y - lo), ok.shape)[ok].astype(np.int64)
    b = np.broadcast_to((hi - lo), ok.shape)[ok].astype(np.int64)
    want = np.minimum(((a.astype(np.float32) / b.astype(np.float32)) * np.float32(256.0)).astype(np.int64), 255)
    m = np.where(b == 1, 0xffffffff, 0xffffffff // b + 1).astype(np.uint64)
    got = np.minimum(((a.astype(np.uint64) << np.uint64(8)) * m) >> np.uint64(32), np.uint64(255)).astype(np.int64)
    assert a.size > 2_000_000 and np.array_equal(got, want)

