"""Parity of the CUDA path (through the C ABI) against the oracle and the golden vectors.  Needs a B200.

Tolerances (BASELINE.json north_star): MFCC images and heat maps max abs error <= 1e-4 against the
reference's float32 output; integer counts (I, U, pos, num) bit-exact whenever the masks agree, with
every disagreeing boundary pixel reported; AUC <= 1e-12.
"""
import numpy as np
import pytest

import acoustic_image_generation_b200 as aig
from acoustic_image_generation_b200 import synth, tables
from oracle import acoustic_oracle as oracle

pytestmark = pytest.mark.gpu

MFCC_TOL = 1e-4          # stated tolerance of the north star
HEAT_TOL = 1e-4
ENERGY_RTOL = 1e-13      # float64 energy: exp() and BLAS-order last-ulp differences only
REF_THR = list(oracle.REFERENCE_THRESHOLDS)
NUM_VARIANTS = 10


@pytest.fixture(scope='module')
def path():
    p = aig.AcousticPath(0)
    assert p.tables_are_reference
    yield p
    p.close()


@pytest.fixture(scope='module')
def torch():
    import torch
    assert torch.cuda.is_available()
    return torch


# ----------------------------------------------------------------------------------------------
# stage 1: MFCC
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize('kind,n', [('chi2', 2), ('lognormal', 1), ('floor', 1)])
def test_mfcc_matches_golden_and_oracle(path, golden, kind, n):
    g = golden('mfcc')
    power = synth.power_frames(n, int(g['seed_' + kind]), kind)
    got = path.mfcc_image(power)
    assert isinstance(got, np.ndarray) and got.dtype == np.float32 and got.shape == (n, 36, 48, 12)
    err = np.abs(got.astype(np.float64) - g['mfcc_' + kind].astype(np.float64)).max()
    print('mfcc %s: max abs err vs reference float32 output = %.3e (max |mfcc| %.1f)' % (kind, err, np.abs(g['mfcc_' + kind]).max()))
    assert err <= MFCC_TOL
    assert np.array_equal(oracle.mfcc_image(power), g['mfcc_' + kind])


@pytest.mark.parametrize('variant', range(NUM_VARIANTS))
def test_mfcc_every_pipeline_variant(path, variant):
    power = synth.power_frames(3, 7, 'chi2')
    want = oracle.mfcc_image(power)
    path.set_mfcc_variant(variant)
    try:
        got = path.mfcc_image(power)
    finally:
        path.set_mfcc_variant(-1)
    assert np.abs(got - want).max() <= MFCC_TOL


def test_mfcc_variants_agree_bitwise(path):
    """Every ring geometry runs the same per-spectrum program: results are identical bit for bit."""
    power = synth.power_frames(2, 8, 'lognormal')
    outs = []
    for variant in range(NUM_VARIANTS):
        path.set_mfcc_variant(variant)
        outs.append(path.mfcc_image(power))
    path.set_mfcc_variant(-1)
    for o in outs[1:]:
        assert np.array_equal(o, outs[0])


def test_mfcc_device_tensor_in_device_tensor_out(path, torch):
    power = synth.power_frames(2, 0, 'chi2')
    d = torch.from_numpy(power).cuda()
    got = path.mfcc_image(d)
    assert got.is_cuda and got.dtype == torch.float32 and tuple(got.shape) == (2, 36, 48, 12)
    assert np.array_equal(got.cpu().numpy(), path.mfcc_image(power))


def test_mfcc_flip180(path):
    power = synth.power_frames(3, 5, 'chi2')
    plain = path.mfcc_image(power)
    flipped = path.mfcc_image(power, flip=True)
    assert np.array_equal(flipped, oracle.flip180(plain))
    assert np.abs(flipped - oracle.mfcc_image(power, flip=True)).max() <= MFCC_TOL


@pytest.mark.parametrize('n_rows', [0, 1, 31, 127, 128, 129, 1727, 1729, 5000])
def test_mfcc_ragged_row_counts(path, n_rows):
    rng = np.random.default_rng(n_rows)
    beam = rng.standard_normal((n_rows, 512), dtype=np.float32) ** 2
    got = path.mfcc_rows(beam)
    assert got.shape == (n_rows, 12)
    if n_rows:
        bank, dct, lifter, mfnorm = oracle.reference_tables()
        want = np.float32(oracle.get_feats(512, beam, 12, dct, mfnorm, lifter, bank))
        assert np.abs(got - want).max() <= MFCC_TOL


def test_mfcc_nan_inf_rows_are_zeroed_like_the_reference(path):
    bank, dct, lifter, mfnorm = oracle.reference_tables()
    beam = np.ones((8, 512), np.float32)
    beam[0, 100] = np.nan
    beam[1, 200] = np.inf
    beam[2, 0] = np.inf        # bin 0 has an all-zero filter row: inf * 0 = NaN in the dense product
    beam[3, 511] = np.nan      # same for the last bin
    beam[4, 300] = -np.inf
    beam[5, 5] = np.inf        # a triangle peak
    with np.errstate(invalid='ignore'):
        want = np.float32(oracle.get_feats(512, beam, 12, dct, mfnorm, lifter, bank))
    got = path.mfcc_rows(beam)
    assert np.isfinite(got).all()
    for r in range(6):
        assert np.array_equal(got[r], np.zeros(12, np.float32)), r
        assert np.array_equal(want[r], np.zeros(12, np.float32)), r
    assert np.abs(got[6:] - want[6:]).max() <= MFCC_TOL and np.abs(got[6]).max() > 0


def test_mfcc_negative_and_tiny_inputs_hit_the_floor(path):
    bank, dct, lifter, mfnorm = oracle.reference_tables()
    rng = np.random.default_rng(3)
    beam = rng.standard_normal((256, 512), dtype=np.float32) * np.float32(1e-3)     # signed: many mel sums < 0.001
    want = np.float32(oracle.get_feats(512, beam, 12, dct, mfnorm, lifter, bank))
    assert np.abs(path.mfcc_rows(beam) - want).max() <= MFCC_TOL


def test_get_feats_dropin_reference_and_generic_tables(path):
    bank, dct, lifter, mfnorm = tables.reference_tables()
    beam = synth.power_frames(1, 9, 'chi2').reshape(-1, 512)[:300]
    got = aig.get_feats(512, beam, 12, dct, mfnorm, lifter, bank)
    assert got.dtype == np.float64 and got.shape == (300, 12)
    assert np.abs(got - oracle.get_feats(512, beam, 12, dct, mfnorm, lifter, bank)).max() <= MFCC_TOL
    # a geometry the fused kernel does not cover: 256 bins, 20 filters, 10 coefficients (generic float64 kernel)
    bank2 = aig.createfilters(256, 20, 300, 4000, 8000)
    dct2, lifter2, mfnorm2 = tables.mfcc_constants(20, 10, 22)
    rng = np.random.default_rng(4)
    beam2 = rng.standard_normal((777, 256), dtype=np.float32) ** 2
    p2 = aig.AcousticPath(0, tables_=(bank2, dct2, lifter2, mfnorm2))
    assert not p2.tables_are_reference
    got2 = p2.mfcc_rows(beam2)
    want2 = oracle.get_feats(256, beam2, 10, dct2, mfnorm2, lifter2, bank2)
    assert np.abs(got2 - np.float32(want2)).max() <= 1e-5
    got3 = aig.get_feats(256, beam2, 10, dct2, mfnorm2, lifter2, bank2)
    assert np.array_equal(got3, got2.astype(np.float64))
    p2.close()


def test_mfcc_host_staging_across_chunks(path):
    """A host batch larger than one 64 MiB staging chunk goes through the double-buffered H2D pipeline."""
    power = synth.power_frames(45, 11, 'chi2')          # 159 MB
    got = path.mfcc_image(power, flip=True)
    want = oracle.mfcc_image(power, flip=True)
    assert np.abs(got - want).max() <= MFCC_TOL


def test_mfcc_large_batch_is_the_small_case_tiled(path, torch):
    """Size-independent property at bench scale: frames are independent, so 2048 frames made of 8 distinct
    frames repeated give exactly the 8-frame answer repeated (7.2 GB of spectra, larger than L2)."""
    base = torch.from_numpy(synth.power_frames(8, 12, 'chi2')).cuda()
    small = path.mfcc_image(base)
    assert np.abs(small.cpu().numpy() - oracle.mfcc_image(base.cpu().numpy())).max() <= MFCC_TOL
    big = base.repeat(256, 1, 1, 1)
    out = path.mfcc_image(big)
    assert torch.equal(out.view(256, 8, 36, 48, 12), small.unsqueeze(0).expand(256, -1, -1, -1, -1))


# ----------------------------------------------------------------------------------------------
# stage 2: normalise, energy, mask, heat map
# ----------------------------------------------------------------------------------------------
def test_normalize_bit_exact(path, golden):
    mf = golden('mfcc')['mfcc_chi2']
    got = path.normalize_images(mf)
    assert np.array_equal(got, golden('energy')['normed_input'])
    assert np.array_equal(got, oracle.normalize_acoustic_images(mf))
    one = aig._normalize_acoustic_images_rescaled(mf[0])
    assert one.shape == (36, 48, 12) and np.array_equal(one, got[0])
    tup = aig._map_func_acoustic_images(mf, 'a', 'v', 1, 2, 'f')
    assert np.array_equal(tup[0], got) and tup[1:] == ('a', 'v', 1, 2, 'f')
    const = path.normalize_images(np.full((1, 36, 48, 12), 3.0, np.float32))
    assert np.isnan(const).all()                          # 0/0, as in the reference


def _mask_report(name, energy, mean, got_mask, want_mask):
    diff = np.argwhere(got_mask != want_mask)
    for idx in diff[:20]:
        f = idx[0]
        print('boundary pixel %s %s: energy %.17g mean %.17g' % (name, tuple(idx), energy[tuple(idx)], mean[f]))
    return len(diff)


@pytest.mark.parametrize('name', ['normed', 'sigmoid', 'smooth'])
def test_energy_and_mask_match_reference(path, golden, name):
    g = golden('energy')
    imgs = {'normed': g['normed_input'], 'sigmoid': synth.sigmoid_images(2, 3), 'smooth': synth.smooth_images(2, 4)}[name]
    energy, mask, scaled, mean = path.energy(imgs, normalize_first=False, want_scaled=True, want_mean=True)
    want = g['energy_' + name]
    assert energy.dtype == np.float64 and mask.dtype == np.uint8
    rel = np.abs(energy - want).max() / np.abs(want).max()
    print('energy %s: max rel err %.3e; bit-identical %.1f%%' % (name, rel, 100 * (energy == want).mean()))
    assert rel <= ENERGY_RTOL
    assert np.abs(mean - g['mean_' + name]).max() <= ENERGY_RTOL * np.abs(g['mean_' + name]).max()
    assert _mask_report(name, energy, mean, mask, g['mask_' + name]) == 0
    if name == 'sigmoid':
        assert np.array_equal(scaled[0], g['sigmoid0_after_call'])    # the in-place side effect, bit-exact


def test_energy_mean_is_numpy_pairwise(path):
    """Fed back its own energies, the device mean equals np.mean bit for bit (same summation tree)."""
    imgs = synth.sigmoid_images(6, 30)
    energy, mask, mean = path.energy(imgs, want_mean=True)
    assert np.array_equal(mean, np.array([np.mean(e) for e in energy]))
    assert np.array_equal(mask, np.stack([oracle.mean_mask(e) for e in energy], 0))


def test_energy_with_normalisation_first(path, golden):
    mf = golden('mfcc')['mfcc_chi2']
    energy, mask = path.energy(mf, normalize_first=True)
    want_e, want_m = oracle.energy_stage(mf, normalize_first=True)
    assert np.abs(energy - want_e).max() <= ENERGY_RTOL * np.abs(want_e).max()
    assert (mask != want_m).sum() == 0


def test_find_logen_dropin_scales_argument_in_place(path, golden, torch):
    g = golden('energy')
    frame = synth.sigmoid_images(2, 3)[0].copy()
    en = aig.find_logen(frame)
    assert en.shape == (36, 48) and en.dtype == np.float64
    assert np.array_equal(frame, g['sigmoid0_after_call'])
    assert np.abs(en - g['energy_sigmoid'][0]).max() <= ENERGY_RTOL * np.abs(en).max()
    keep = synth.sigmoid_images(2, 3)[0].copy()
    path.find_logen(keep, inplace=False)
    assert np.array_equal(keep, synth.sigmoid_images(2, 3)[0])
    dev = torch.from_numpy(synth.sigmoid_images(2, 3)[0].copy()).cuda()
    en_dev = path.find_logen(dev)
    assert en_dev.is_cuda and np.array_equal(dev.cpu().numpy(), g['sigmoid0_after_call'])
    assert np.array_equal(en_dev.cpu().numpy(), en)


@pytest.mark.parametrize('exact', [0, 1])
@pytest.mark.parametrize('shape', [(224, 298), (224, 224), (36, 48), (100, 77), (448, 596)])
def test_heatmap_matches_oracle_and_cv2_golden(path, golden, shape, exact):
    """Default: float32 bilinear on the frame-normalised map (HBM-write-bound, error ~1e-7); heatmap_exact = 1: the
    float64 replica of the oracle's arithmetic (error exactly 0)."""
    e = golden('energy')['energy_smooth']
    path.set_option('heatmap_exact', exact)
    try:
        got = path.heatmap(e, *shape)
        flat = path.heatmap(np.full((1, 36, 48), 0.25), *shape)
    finally:
        path.set_option('heatmap_exact', 0)
    want = np.stack([oracle.heatmap(x, *shape) for x in e], 0)
    assert got.dtype == np.float32 and got.shape == (2,) + shape
    err = np.abs(got.astype(np.float64) - want).max()
    print('heatmap %s exact=%d: max abs err %.3e' % (shape, exact, err))
    assert err <= (0.0 if exact else 2e-6) and err <= HEAT_TOL
    assert got.min() == 0.0 and abs(float(got.max()) - 1.0) <= 1e-6
    assert np.isnan(flat).all()                              # constant map: (x - min) / (max - min) = 0/0, as matplotlib
    # nearly flat map (range 1e-9 of the level): float32 on raw energies would lose it, the normalised form does not
    base = np.full((1, 36, 48), 0.04)
    base[0, 10:20, 5:30] += 4e-11
    path.set_option('heatmap_exact', exact)
    try:
        tiny = path.heatmap(base, *shape)
    finally:
        path.set_option('heatmap_exact', 0)
    assert np.abs(tiny[0] - oracle.heatmap(base[0], *shape)).max() <= 2e-6
    key = 'up_%d_%d' % shape
    if key in golden('heatmap'):
        up = golden('heatmap')[key]                       # cv2.resize output of frame 0
        ref = np.float32((up - up.min()) / (up.max() - up.min()))
        assert np.abs(got[0] - ref).max() <= HEAT_TOL


@pytest.mark.parametrize('shape', [(224, 298), (20, 30), (500, 37), (2048, 5), (37, 49), (3, 2048), (2048, 2048), (1, 1)])
def test_heatmap_extremes_on_rough_maps(path, shape):
    """The fast kernel looks for the image's min / max only on the first and last output row of each source-row pair
    (bilinear interpolation is monotonic in between): on rough maps, where extremes sit anywhere, the result must still
    span exactly [0, ~1] and match the oracle."""
    rng = np.random.default_rng(shape[0] * 7 + shape[1])
    e = rng.random((6, 36, 48)) ** 4
    e[1, 35, :] += 3.0          # extremes on the clamped border rows
    e[2, 0, 0] = -2.0
    got = path.heatmap(e, *shape)
    for i in range(6):
        want = oracle.heatmap(e[i], *shape)
        if shape == (1, 1):                                      # one output pixel: (x - x) / (x - x) = NaN, like the reference
            assert np.isnan(got[i]).all() and np.isnan(want).all()
            continue
        assert np.abs(got[i] - want).max() <= 2e-6
        assert got[i].min() == 0.0 and abs(float(got[i].max()) - 1.0) <= 1e-6


@pytest.mark.parametrize('shape', [(224, 298), (224, 224), (36, 48), (73, 95)])
def test_resize_mask_bit_exact(path, golden, shape):
    e = golden('energy')
    masks = np.concatenate([e['mask_smooth'], e['mask_sigmoid']], 0)
    got = path.resize_mask(masks, *shape)
    want = np.stack([oracle.resize_mask(m, *shape) for m in masks], 0)
    assert np.array_equal(got, want)
    if shape == (224, 298):
        cvg = np.unpackbits(golden('heatmap')['mask_up_224_298'], axis=-1)[..., :298]
        assert np.array_equal(got[:2], cvg)


# ----------------------------------------------------------------------------------------------
# stage 3: IoU sweeps and AUC
# ----------------------------------------------------------------------------------------------
def test_acivw_iou_sweep_bit_exact(path, golden):
    g = golden('acivw_iou')
    n = int(g['num'])
    a = synth.smooth_images(n, 10)
    b = synth.smooth_images(n, 11)
    b[: n // 2] = a[: n // 2] * np.float32(0.9) + b[: n // 2] * np.float32(0.1)
    _, mask_a = path.energy(a)
    _, mask_b = path.energy(b)
    for thr, key in ((REF_THR, 'pos11'), (np.linspace(0, 1, 101), 'pos101')):
        inter, union, pos, num = path.iou_sweep(mask_a, mask_b, thr)
        assert np.array_equal(inter, g['inter']) and np.array_equal(union, g['union'])
        assert np.array_equal(pos, g[key]) and num == n
    # accumulate over two shards == one pass (what the multi-GPU all-reduce relies on)
    _, _, pos1, num1 = path.iou_sweep(mask_a[:20], mask_b[:20], REF_THR)
    _, _, pos2, num2 = path.iou_sweep(mask_a[20:], mask_b[20:], REF_THR, pos=pos1, num=num1)
    assert np.array_equal(pos2, g['pos11']) and num2 == n
    rates = aig.success_rates(pos2, num2)
    assert abs(aig.auc(REF_THR, rates) - float(golden('auc')['acivw11'])) <= 1e-12


def test_iou_sweep_edge_cases(path):
    empty = np.zeros((3, 36, 48), np.uint8)
    full = np.ones((3, 36, 48), np.uint8)
    inter, union, pos, num = path.iou_sweep(empty, empty, [0.0, 0.5])
    assert inter.tolist() == [0, 0, 0] and union.tolist() == [0, 0, 0]
    assert pos.tolist() == [0, 0] and num == 3                     # 0/0 = NaN never counts, but num does
    inter, union, pos, num = path.iou_sweep(full * 7, full, [0.0, 0.5, 1.0])   # any non-zero byte is "set"
    assert inter.tolist() == [1728] * 3 and pos.tolist() == [3, 3, 0]
    rng = np.random.default_rng(5)
    a = (rng.random((257, 36, 48)) > 0.5).astype(np.uint8)
    b = (rng.random((257, 36, 48)) > 0.7).astype(np.uint8)
    thr = np.linspace(0, 1, 1024)
    inter, union, pos, num = path.iou_sweep(a, b, thr)
    want = [oracle.iou_pair(x, y) for x, y in zip(a, b)]
    assert inter.tolist() == [w[0] for w in want] and union.tolist() == [w[1] for w in want]
    wpos, wnum = oracle.success_counts([w[2] for w in want], thr)
    assert np.array_equal(pos, wpos) and num == wnum


@pytest.mark.parametrize('shape', [(224, 298), (224, 224)])
def test_flickr_ciou_sweep_bit_exact(path, golden, shape):
    g = golden('flickr_ciou')
    n = int(g['num'])
    pred = synth.smooth_images(n, 20)
    if shape == (224, 298):
        xmin, xmax, ymin, ymax = synth.flickr_boxes(n, 21)
    else:
        xmin, xmax, ymin, ymax = synth.flickr_boxes(n, 22, 224, 224)
    _, masks = path.energy(pred)
    for thr in (REF_THR, np.linspace(0, 1, 101)):
        i2, u2, pos, num = path.ciou_sweep(masks, xmin, xmax, ymin, ymax, thr, out_hw=shape)
        wi, wu, wpos, wnum = oracle.flickr_sweep(masks, xmin, xmax, ymin, ymax, thr, *shape)
        assert np.array_equal(i2, wi) and np.array_equal(u2, wu)
        assert np.array_equal(pos, wpos) and num == wnum
    if shape == (224, 298):
        i2, u2, pos, num = path.ciou_sweep(masks, xmin, xmax, ymin, ymax, REF_THR)
        assert np.array_equal(i2 / 2.0, g['inter']) and np.array_equal(u2 / 2.0, g['union'])
        assert np.array_equal(pos, g['pos11'])
        assert abs(aig.auc(REF_THR, aig.success_rates(pos, num)) - float(golden('auc')['flickr11'])) <= 1e-12


@pytest.mark.parametrize('shape', [(2048, 2048), (1, 1), (2, 2047), (1500, 3)])
def test_mask_resize_and_ciou_at_extreme_output_sizes(path, shape):
    """The separable integer kernels at the largest geometry the ABI accepts (184 KiB of shared rows) and at degenerate
    ones, against the oracle's integer restatement."""
    rng = np.random.default_rng(shape[0] + shape[1])
    masks = (rng.random((2, 36, 48)) > 0.6).astype(np.uint8)
    got = path.resize_mask(masks, *shape)
    want = np.stack([oracle.resize_mask(m, *shape) for m in masks], 0)
    assert np.array_equal(got, want)
    h, w = shape
    xmin = np.array([[0, w // 3, 0], [w // 2, 0, 0]], np.int32); xmax = np.array([[w // 2, w, 0], [w, 0, 0]], np.int32)
    ymin = np.array([[0, h // 4, 0], [h // 2, 0, 0]], np.int32); ymax = np.array([[h // 2, h, 0], [h, 0, 0]], np.int32)
    i2, u2, pos, num = path.ciou_sweep(masks, xmin, xmax, ymin, ymax, REF_THR, out_hw=shape)
    wi, wu, wpos, wnum = oracle.flickr_sweep(masks, xmin, xmax, ymin, ymax, REF_THR, *shape)
    assert np.array_equal(i2, wi) and np.array_equal(u2, wu) and np.array_equal(pos, wpos) and num == wnum


def test_ciou_box_edge_cases(path):
    mask = np.zeros((4, 36, 48), np.uint8)
    mask[1] = 1
    mask[2, 10:20, 10:30] = 1
    mask[3, :, :24] = 1
    xmin = np.array([[0, 0, 0], [0, 290, 0], [50, 50, 50], [-5, 100, 0]], np.int32)
    xmax = np.array([[0, 0, 0], [10, 298, 0], [200, 200, 200], [400, 20, 0]], np.int32)   # absent / clipped / x3 / swapped
    ymin = np.array([[0, 0, 0], [0, 200, 0], [40, 40, 40], [-3, 10, 0]], np.int32)
    ymax = np.array([[0, 0, 0], [5, 224, 0], [180, 180, 180], [300, 5, 0]], np.int32)
    i2, u2, pos, num = path.ciou_sweep(mask, xmin, xmax, ymin, ymax, REF_THR)
    wi, wu, wpos, wnum = oracle.flickr_sweep(mask, xmin, xmax, ymin, ymax, REF_THR)
    assert np.array_equal(i2, wi) and np.array_equal(u2, wu) and np.array_equal(pos, wpos) and num == wnum
    assert u2[0] == 0 and pos.sum() >= 0


# ----------------------------------------------------------------------------------------------
# chained path
# ----------------------------------------------------------------------------------------------
def test_mfcc_energy_chain_against_oracle_chain(path, torch):
    """End to end from spectra: stage-1 rounding (<= 2e-5) may flip pixels whose energy sits at the mean; every
    such boundary pixel is listed, and counts are exact wherever the masks agree."""
    power = synth.power_frames(6, 40, 'chi2')
    mfcc, energy, mask, mean = path.mfcc_energy(power, flip=True, normalize_first=True, want_mean=True)
    want_mfcc = oracle.mfcc_image(power, flip=True)
    assert np.abs(mfcc - want_mfcc).max() <= MFCC_TOL
    # (i) stage 2 alone, fed the oracle's MFCC image: exact masks
    e2, m2 = path.energy(want_mfcc, normalize_first=True)
    want_e, want_m = oracle.energy_stage(want_mfcc, normalize_first=True)
    assert (m2 != want_m).sum() == 0
    # (ii) end to end
    flips = _mask_report('chain', energy, mean, mask, want_m)
    print('end-to-end mask disagreements: %d of %d pixels' % (flips, mask.size))
    assert flips <= 0.002 * mask.size
    rel = np.abs(energy - want_e).max() / np.abs(want_e).max()
    print('end-to-end energy max rel err %.3e' % rel)
    assert rel <= 1e-3
    # device in / device out gives the same bits as host in / host out
    d = path.mfcc_energy(torch.from_numpy(power).cuda(), flip=True, normalize_first=True)
    assert np.array_equal(d[0].cpu().numpy(), mfcc) and np.array_equal(d[1].cpu().numpy(), energy)
    assert np.array_equal(d[2].cpu().numpy(), mask)


@pytest.mark.parametrize('flip,normalize', [(True, True), (False, False)])
def test_chain_modes_agree_bitwise(path, torch, flip, normalize):
    """The fused persistent kernel (mode 2, every ring geometry), the two-stream overlap (mode 1) and the plain
    sequence (mode 0) run the same arithmetic: identical bits, over enough frames (700 > 4 per CTA) to cycle
    the frame hand-off slots and barrier phases several times."""
    n = 700
    base = torch.from_numpy(synth.power_frames(7, 41, 'chi2')).cuda()
    power = base.repeat(100, 1, 1, 1)
    results = []
    try:
        for mode, variant in ((0, 0), (1, 0), (2, 0), (2, 1), (2, 2)):
            path.set_option('chain_mode', mode)
            path.set_option('fused_variant', variant)
            mfcc, energy, mask, mean = path.mfcc_energy(power, flip=flip, normalize_first=normalize, want_mean=True)
            results.append((mfcc.clone(), energy.clone(), mask.clone(), mean.clone()))
    finally:
        path.set_option('chain_mode', 2)
        path.set_option('fused_variant', 0)
    ref = results[0]
    for got in results[1:]:
        for a, b in zip(ref, got):
            assert torch.equal(a, b)
    # and they are the 7-frame answer tiled (frames are independent)
    small = path.mfcc_energy(base, flip=flip, normalize_first=normalize)
    assert torch.equal(ref[0].view(100, 7, 36, 48, 12), small[0].unsqueeze(0).expand(100, -1, -1, -1, -1))
    assert torch.equal(ref[2].view(100, 7, 36, 48), small[2].unsqueeze(0).expand(100, -1, -1, -1))
    want_mfcc = oracle.mfcc_image(base.cpu().numpy(), flip=flip)
    assert np.abs(small[0].cpu().numpy() - want_mfcc).max() <= MFCC_TOL


@pytest.mark.parametrize('n', [1, 2, 147, 149, 300])
def test_fused_kernel_ragged_frame_counts(path, torch, n):
    base = torch.from_numpy(synth.power_frames(3, 42, 'lognormal')).cuda()
    power = base.repeat((n + 2) // 3, 1, 1, 1)[:n].contiguous()
    path.set_option('chain_mode', 2)
    fused = path.mfcc_energy(power, flip=True, normalize_first=True)
    path.set_option('chain_mode', 0)
    try:
        seq = path.mfcc_energy(power, flip=True, normalize_first=True)
    finally:
        path.set_option('chain_mode', 2)
    for a, b in zip(fused, seq):
        assert torch.equal(a, b)


def test_launch_counter_counts_kernels(path):
    before = path.launch_count
    path.energy(synth.sigmoid_images(1, 0))
    assert path.launch_count == before + 1


# ----------------------------------------------------------------------------------------------
# evaluation drivers (the callers of the scoring path) and their metric files
# ----------------------------------------------------------------------------------------------
def test_acivw_evaluation_driver_writes_reference_files(path, golden, tmp_path):
    from acoustic_image_generation_b200 import evaluate, metrics_io
    g = golden('acivw_iou')
    n = int(g['num'])
    a = synth.smooth_images(n, 10)
    b = synth.smooth_images(n, 11)
    b[: n // 2] = a[: n // 2] * np.float32(0.9) + b[: n // 2] * np.float32(0.1)
    ev = evaluate.AcivwEvaluation(path)
    for lo in range(0, n, 16):                         # batches, like the session.run loop
        ev.add_batch(a[lo:lo + 16], b[lo:lo + 16])
    res = ev.finish(str(tmp_path))
    assert np.array_equal(res['pos'], g['pos11']) and res['num'] == n
    assert abs(res['auc'] - float(golden('auc')['acivw11'])) <= 1e-12
    for t, p in zip(REF_THR, g['pos11']):
        assert metrics_io.read_accuracy_file(str(tmp_path), t) == float('{:6f}'.format(p / n))
    assert open(str(tmp_path / 'area.txt')).read() == 'area {:6f}'.format(res['auc'])


def test_flickr_evaluation_driver(path, golden, torch):
    from acoustic_image_generation_b200 import evaluate
    g = golden('flickr_ciou')
    n = int(g['num'])
    pred = torch.from_numpy(synth.smooth_images(n, 20)).cuda()      # device-resident inputs
    boxes = [torch.from_numpy(v).cuda() for v in synth.flickr_boxes(n, 21)]
    ev = evaluate.FlickrEvaluation(path)
    ev.add_batch(pred[:20], *[v[:20] for v in boxes])
    ev.add_batch(pred[20:], *[v[20:] for v in boxes])
    res = ev.finish()
    assert np.array_equal(res['pos'], g['pos11']) and res['num'] == n
    assert abs(res['auc'] - float(golden('auc')['flickr11'])) <= 1e-12


# ----------------------------------------------------------------------------------------------
# "next" rows N1 (audio front half) and N2 (batch assembly)
# ----------------------------------------------------------------------------------------------
def test_power_spectrum_matches_numpy_rfft(path, golden):
    g = golden('audio_front')
    for audio in (synth.audio_rows(24, 90, np.int32), synth.audio_rows(8, 91, np.float32, amplitude=1.0)):
        got = path.power_spectrum(audio)
        want = oracle.power_spectrum(audio, oracle.tukey_window())
        assert got.dtype == np.float32 and got.shape == want.shape
        # float64 FFT rounded to float32: relative to the row's peak power the error is at float32 rounding level
        rel = np.abs(got - want).max(axis=1) / want.max(axis=1)
        assert rel.max() <= 1e-7
        bare = path.power_spectrum(audio, window=None)
        want_bare = oracle.power_spectrum(audio, None)
        assert (np.abs(bare - want_bare).max(axis=1) / want_bare.max(axis=1)).max() <= 1e-7
    assert np.abs(path.power_spectrum(synth.audio_rows(24, 90, np.int32))[:4] - g['power_int_f64']).max() \
        <= 1e-7 * g['power_int_f64'].max()


def test_build_spectrograms_dropin_matches_reference(path, golden):
    g = golden('audio_front')
    for key, audio in (('mfcc_int', synth.audio_rows(24, 90, np.int32)),
                       ('mfcc_float', synth.audio_rows(8, 91, np.float32, amplitude=1.0))):
        got = aig._build_spectrograms_function(audio)
        assert got.dtype == np.float32 and got.shape == g[key].shape
        err = np.abs(got - g[key]).max()
        print('%s: max abs err %.3e (max |mfcc| %.1f)' % (key, err, np.abs(g[key]).max()))
        assert err <= MFCC_TOL


def test_butter_lowpass_filter_matches_scipy(path, golden, torch):
    g = golden('audio_front')
    audio_i = synth.audio_rows(24, 90, np.int32)
    audio_f = synth.audio_rows(8, 91, np.float32, amplitude=1.0)
    got_i = aig.butter_lowpass_filter(audio_i)
    got_f = path.butter_lowpass_filter(audio_f)
    assert got_i.dtype == np.float32 and got_i.shape == audio_i.shape
    # same float64 recurrence in the same order: agreement far below the float32 output resolution
    for got, key in ((got_i, 'lowpass_int'), (got_f, 'lowpass_float')):
        scale = np.abs(g[key]).max()
        err = np.abs(got.astype(np.float64) - g[key].astype(np.float64)).max() / scale
        print('%s: max err / max |y| = %.3e, bit-identical %.1f%%' % (key, err, 100 * (got == g[key]).mean()))
        assert err <= 1e-6
    dev = path.butter_lowpass_filter(torch.from_numpy(audio_i).cuda())
    assert dev.is_cuda and np.array_equal(dev.cpu().numpy(), got_i)
    # low-passed audio through the spectrogram function, as the dataloader chains them (:78-82)
    mf = aig._build_spectrograms_function(got_i)
    assert np.abs(mf - g['mfcc_lowpassed_int']).max() <= 5e-4     # |MFCC| up to ~100 on a 1e9-range spectrum


def test_normalize_and_tile_mfcc_bit_exact(path, torch):
    rng = np.random.default_rng(1)
    m = (rng.standard_normal((301, 12)) * 7).astype(np.float32)
    assert np.array_equal(path.normalize_mfcc(m), oracle.normalize_mfcc(m))
    assert np.array_equal(aig._normalize_mfcc(m[5]), oracle.normalize_mfcc(m[5:6])[0])
    tup = aig._map_func_mfcc('img', m, 'vid', 1, 2, m[::-1].copy())
    assert np.array_equal(tup[1], oracle.normalize_mfcc(m)) and np.array_equal(tup[5], oracle.normalize_mfcc(m[::-1]))
    assert tup[0] == 'img' and tup[2:5] == ('vid', 1, 2)
    assert np.array_equal(path.tile_mfcc(m), oracle.tile_mfcc(m))
    assert np.array_equal(path.tile_mfcc(m, normalize=True), oracle.tile_mfcc(oracle.normalize_mfcc(m)))
    dev = path.tile_mfcc(torch.from_numpy(m).cuda())
    assert dev.is_cuda and tuple(dev.shape) == (301, 36, 48, 12)
    assert np.isnan(path.normalize_mfcc(np.full((1, 12), 2.0, np.float32))).all()


def test_triplet_slices_and_losses(path, torch):
    """trainer/mfcctrainer.py:103-117: the four channel triplets and the five mean-squared-error terms."""
    for n in (1, 7, 130):
        a, b = synth.sigmoid_images(n, 3), synth.sigmoid_images(n, 4)
        got = path.split_triplets(a)
        assert got.shape == (4, n, 36, 48, 3) and np.array_equal(got, oracle.split_triplets(a))
        mse = path.triplet_mse(a, b)
        want = oracle.triplet_mse(a, b)
        assert mse.shape == (5,) and np.abs(mse / want - 1).max() <= 1e-13
        assert np.array_equal(mse, path.triplet_mse(a, b))                    # fixed summation order
        assert abs(mse[0] - np.mean(mse[1:])) <= 1e-15                        # equal-sized slices
    assert np.array_equal(path.triplet_mse(a, a), np.zeros(5))
    dev = path.split_triplets(torch.from_numpy(a).cuda())
    assert dev.is_cuda and np.array_equal(dev.cpu().numpy(), oracle.split_triplets(a))
    assert np.array_equal(path.triplet_mse(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()), mse)
    with pytest.raises(ValueError):
        path.triplet_mse(a, b[:-1])


def test_overlay_bit_exact_vs_oracle(path, golden, torch):
    lut = tables.jet_lut()
    e = golden('energy')['energy_smooth']
    heat = path.heatmap(e)                                   # [2, 224, 298] float32
    rng = np.random.default_rng(2)
    frames = rng.integers(0, 256, (2, 224, 298, 3), dtype=np.uint8)
    frames[1, :, :, :] = 77                                  # constant frame: zero gray range
    got = path.overlay(heat, frames)
    assert got.dtype == np.uint8 and got.shape == (2, 224, 298, 3)
    for i in range(2):
        assert np.array_equal(got[i], oracle.overlay(heat[i], frames[i], lut))
    bare = path.overlay(heat)
    assert np.array_equal(bare[0], oracle.overlay(heat[0], None, lut))
    # odd geometry takes the scalar kernel: 15 x 17 = 255 pixels per frame, heat values on and past both ends
    small = rng.random((3, 15, 17)).astype(np.float32)
    small[0, 0, :4] = [0.0, 1.0, 1.5, -0.25]
    small_frames = rng.integers(0, 256, (3, 15, 17, 3), dtype=np.uint8)
    got = path.overlay(small, small_frames)
    for i in range(3):
        assert np.array_equal(got[i], oracle.overlay(small[i], small_frames[i], lut))
    assert np.array_equal(path.overlay(small)[1], oracle.overlay(small[1], None, lut))
    dev = path.overlay(torch.from_numpy(heat).cuda(), torch.from_numpy(frames).cuda(), alpha=0.5)
    assert dev.is_cuda and np.array_equal(dev.cpu().numpy()[0], oracle.overlay(heat[0], frames[0], lut, 0.5))


def test_overlay_luma_plane_form_equals_two_read_form(path, torch):
    """aig_overlay keeps a frame's luma plane in shared memory between its passes (option overlay_luma, default on); the
    form that reads the BGR frame in both passes serves larger frames.  Both against the oracle and against each other,
    on more frames than CTAs in flight (a CTA's plane is overwritten by its next frame), with constant and two-level
    frames, at a size just inside the shared-memory limit and one past it."""
    lut = tables.jet_lut()
    rng = np.random.default_rng(11)
    for (n, h, w) in ((700, 24, 36), (3, 224, 298), (5, 30, 34), (2, 256, 320), (2, 256, 324)):     # 81 920 px = the limit; 82 944 px falls back
        heat = torch.rand(n, h, w, device='cuda')
        frames = torch.from_numpy(rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)).cuda()
        frames[0] = 200
        frames[1, : h // 2] = 3
        frames[1, h // 2:] = 250
        try:
            path.set_option('overlay_luma', 1)
            a = path.overlay(heat, frames)
            path.set_option('overlay_luma', 0)
            b = path.overlay(heat, frames)
        finally:
            path.set_option('overlay_luma', 1)
        assert torch.equal(a, b)
        hn, fn, an = heat.cpu().numpy(), frames.cpu().numpy(), a.cpu().numpy()
        for i in sorted({0, 1, n - 1, n // 2}):
            assert np.array_equal(an[i], oracle.overlay(hn[i], fn[i], lut))


# ----------------------------------------------------------------------------------------------
# host staging: ordinary (pageable) NumPy inputs go up through the pinned ring of host_staging.h
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize('threads', [-1, 0, 1, 3])
def test_pageable_inputs_match_device_inputs(path, torch, threads):
    """Same bits whether the spectra arrive as device tensors, pinned arrays or pageable arrays staged by 0..n copy
    threads; sizes below, at and above the 8 MiB staging threshold, the 4 MiB slot size and the 64 MiB device chunk."""
    path.set_option('host_copy_threads', threads)
    try:
        for n in (1, 3, 19, 40):
            power = synth.power_frames(n, 5, 'chi2')                     # pageable
            want = [t.cpu().numpy() for t in path.mfcc_energy(torch.from_numpy(power).cuda(), flip=True)]
            for _ in range(2):                                           # second call reuses ring slots still in flight
                got = path.mfcc_energy(power, flip=True)
                for a, b in zip(got, want):
                    assert np.array_equal(a, b, equal_nan=True)
            rows = path.mfcc_rows(power.reshape(-1, 512)[:-7])           # ragged row count through aig_mfcc
            assert np.array_equal(rows, oracle_free_rows(path, torch, power.reshape(-1, 512)[:-7]))
        big = synth.power_frames(110, 6, 'chi2')                         # 9.1 MB of MFCC come back through the ring too
        want = [t.cpu().numpy() for t in path.mfcc_energy(torch.from_numpy(big).cuda(), flip=True)]
        for a, b in zip(path.mfcc_energy(big, flip=True), want):
            assert np.array_equal(a, b, equal_nan=True)
        energies = np.random.default_rng(3).random((70, 36, 48))         # 18.7 MB of heat maps into a pageable array
        assert np.array_equal(path.heatmap(energies), path.heatmap(torch.from_numpy(energies).cuda()).cpu().numpy())
        images = synth.sigmoid_images(300, 2)                            # 24.9 MB through the generic Io staging
        e_dev, m_dev = path.energy(torch.from_numpy(images).cuda())
        e, m = path.energy(images)
        assert np.array_equal(e, e_dev.cpu().numpy()) and np.array_equal(m, m_dev.cpu().numpy())
    finally:
        path.set_option('host_copy_threads', -1)


@pytest.mark.parametrize('streaming', [0, 1])
def test_staging_fill_forms_agree(path, torch, streaming):
    """The pinned slots filled by memcpy or by non-temporal stores (host_copy.cpp): same bits as a device-resident pass,
    for sources at every 4-byte alignment within a 32-byte vector and ragged sizes; bad option values are refused."""
    rows = synth.power_frames(12, 21, 'chi2').reshape(-1, 512)                  # 42 MB
    full = path.mfcc_rows(torch.from_numpy(rows).cuda()).cpu().numpy()
    raw = np.empty(rows.nbytes + 64, np.uint8)
    path.set_option('host_copy_streaming', streaming)
    try:
        for shift in (0, 4, 12, 28):
            base = (-raw.ctypes.data) % 32 + shift                              # source address = shift (mod 32)
            view = raw[base:base + rows.nbytes].view(np.float32).reshape(rows.shape)
            view[...] = rows
            for r in (rows.shape[0], rows.shape[0] - 1, 4099):
                assert np.array_equal(path.mfcc_rows(view[:r]), full[:r], equal_nan=True), (shift, r)
        with pytest.raises(Exception):
            path.set_option('host_copy_streaming', 2)
    finally:
        path.set_option('host_copy_streaming', -1)


@pytest.mark.parametrize('piece', [64 << 10, 512 << 10, 4 << 20])
def test_small_staged_uploads(path, torch, piece):
    """Mid-size pageable arrays (the reference's evaluation batches: 16 images = 1.3 MB) through the staging ring in
    small pieces (options staged_min_bytes / staged_solo_bytes: by the copy threads or by the calling thread alone), around
    the threshold and with ragged last pieces: same bits as device inputs."""
    path.set_option('staged_small_piece_bytes', piece)
    try:
        for min_bytes, solo in ((65536, 0), (1 << 20, 0), (1 << 20, 4 << 20), (8 << 20, 0)):
            path.set_option('staged_min_bytes', min_bytes)
            path.set_option('staged_solo_bytes', solo)
            for n in (1, 3, 13, 16, 49):
                real, recon = synth.sigmoid_images(n, 40 + n), synth.sigmoid_images(n, 80 + n)
                want = path.acivw_batch(torch.from_numpy(real).cuda(), torch.from_numpy(recon).cuda())
                as_np = lambda x: x.cpu().numpy() if hasattr(x, 'cpu') else np.asarray(x)
                for _ in range(2):
                    got = path.acivw_batch(real, recon)
                    for g, w in zip(got, want):
                        assert np.array_equal(as_np(g), as_np(w))
                e_dev, m_dev = path.energy(torch.from_numpy(real).cuda())
                e, m = path.energy(real)
                assert np.array_equal(e, e_dev.cpu().numpy()) and np.array_equal(m, m_dev.cpu().numpy())
        with pytest.raises(Exception):
            path.set_option('staged_min_bytes', 1000)
    finally:
        path.set_option('staged_min_bytes', 8 << 20)
        path.set_option('staged_solo_bytes', 0)
        path.set_option('staged_small_piece_bytes', 512 << 10)


def test_staging_ring_soak_over_random_sizes(path, torch):
    """Thirty uploads of random sizes back to back (8 - 250 MB, ragged last pieces, ring slots of one call still in
    flight when the next starts), results compared with one device-resident pass over the same rows."""
    rng = np.random.default_rng(99)
    rows = synth.power_frames(70, 8, 'lognormal').reshape(-1, 512)              # 248 MB pageable
    full = path.mfcc_rows(torch.from_numpy(rows).cuda()).cpu().numpy()
    for _ in range(30):
        r = int(rng.integers(4096, rows.shape[0] + 1))
        got = path.mfcc_rows(rows[:r])
        assert np.array_equal(got, full[:r], equal_nan=True), r
    big_out = path.tile_mfcc(full[:20000])                                      # 1.66 GB result drained through the ring
    assert big_out.shape == (20000, 36, 48, 12) and np.array_equal(big_out[123, 5, 7], full[123])
    assert np.array_equal(big_out[19999, 35, 47], full[19999]) and np.array_equal(big_out[:, 0, 0, :], full[:20000])


def oracle_free_rows(path, torch, rows):
    """aig_mfcc on a device copy of the same rows (no host staging involved)."""
    return path.mfcc_rows(torch.from_numpy(np.ascontiguousarray(rows)).cuda()).cpu().numpy()


# ----------------------------------------------------------------------------------------------
# error behaviour of the C ABI (negative status + message, never a crash, never a CPU fallback)
# ----------------------------------------------------------------------------------------------
def test_argument_errors_are_reported(path):
    import ctypes
    from acoustic_image_generation_b200 import _lib
    lib = _lib.load()
    h = path._h
    buf = np.zeros((1728, 512), np.float32)
    out = np.zeros((1728, 12), np.float32)
    # flip180 needs whole frames
    assert lib.aig_mfcc(h, buf.ctypes.data, 1000, out.ctypes.data, 1, 1728) == -1
    assert b'multiple of frame_pixels' in lib.aig_last_error(h)
    # null buffers
    assert lib.aig_mfcc(h, None, 10, out.ctypes.data, 0, 1) == -1
    assert lib.aig_energy(h, None, 1, 0, None, None, None, None) == -1
    # negative counts, oversize outputs, too many thresholds
    assert lib.aig_normalize_images(h, buf.ctypes.data, -1, buf.ctypes.data) == -1
    e = np.zeros((1, 36, 48)); big = np.zeros(4, np.float32)
    assert lib.aig_heatmap(h, e.ctypes.data, 1, 4096, 10, big.ctypes.data) == -1
    thr = np.linspace(0, 1, 2000); pos = np.zeros(2000, np.int64); num = np.zeros(1, np.int64)
    m = np.zeros((1, 1728), np.uint8)
    assert lib.aig_iou_sweep(h, m.ctypes.data, m.ctypes.data, 1, thr.ctypes.data, 2000, None, None, pos.ctypes.data, num.ctypes.data) == -1
    assert lib.aig_set_option(h, b'no_such_option', 1) == -1
    assert lib.aig_set_option(h, b'mfcc_variant', 99) == -1
    with pytest.raises(aig.AigError):
        path.set_option('chain_mode', 7)
    with pytest.raises(ValueError):
        path.mfcc_image(np.zeros((1, 36, 48, 100), np.float32))
    with pytest.raises(aig.AigError):
        aig.auc([0.0], [1.0])
    # filtfilt: rows must be longer than the padding
    b, a, zi = tables.butter_lowpass()
    with pytest.raises(aig.AigError):
        path.butter_lowpass_filter(np.zeros((2, 20), np.float32))
    # the chained pass refuses non-reference tables instead of silently using the wrong filter bank
    bank2 = aig.createfilters(256, 20, 300, 4000, 8000)
    dct2, lifter2, mfnorm2 = tables.mfcc_constants(20, 10, 22)
    p2 = aig.AcousticPath(0, tables_=(bank2, dct2, lifter2, mfnorm2))
    with pytest.raises(aig.AigError) as info:
        p2.mfcc_energy(np.zeros((1, 36, 48, 512), np.float32))
    assert info.value.code == -3
    p2.close()
    # zero-sized work is a no-op
    assert path.mfcc_rows(np.zeros((0, 512), np.float32)).shape == (0, 12)
    assert lib.aig_energy(h, buf.ctypes.data, 0, 0, None, None, None, None) == 0


def test_four_worker_threads_each_with_their_own_handle(golden):
    """The reference's tf.data map runs num_parallel_calls=4 Python workers (outdoor_data_mfcc.py:82) that call
    get_feats / find_logen concurrently; the drop-ins keep one handle per thread and ctypes releases the GIL."""
    import threading
    bank, dct, lifter, mfnorm = tables.reference_tables()
    power = synth.power_frames(4, 33, 'chi2').reshape(4, -1, 512)
    want = [oracle.get_feats(512, p, 12, dct, mfnorm, lifter, bank) for p in power]
    imgs = synth.sigmoid_images(4, 34)
    want_e = [oracle.find_logen(f.copy()) for f in imgs]
    results, errors, handles = [None] * 4, [], [None] * 4
    # 14 MB per call: above the staging threshold, so four handles run their pinned rings and copy threads at once
    batches = [synth.power_frames(4, 40 + i, 'chi2') for i in range(4)]
    main_path = aig.default_path()
    want_batch = [main_path.mfcc_image(b, flip=True) for b in batches]
    staged = [None] * 4

    def worker(i):
        try:
            for _ in range(5):
                feats = aig.get_feats(512, power[i], 12, dct, mfnorm, lifter, bank)
                en = aig.find_logen(imgs[i].copy())
                staged[i] = aig.default_path().mfcc_image(batches[i], flip=True)
            results[i] = (feats, en)
            handles[i] = id(aig.default_path())
        except Exception as exc:            # pragma: no cover
            errors.append(exc)

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    assert len(set(handles)) == 4                       # one libaig handle per worker thread
    for i in range(4):
        assert np.abs(results[i][0] - want[i]).max() <= MFCC_TOL
        assert np.abs(results[i][1] - want_e[i]).max() <= ENERGY_RTOL * np.abs(want_e[i]).max()
        assert np.array_equal(staged[i], want_batch[i])


def test_record_batches_on_the_device_match_the_oracle_chain(tmp_path, torch):
    """loader.AcousticBatches: TFRecords -> flipped, normalised acoustic images + normalised audio MFCC + mfccmap +
    low-passed waveform as CUDA tensors, against the oracle applied to the same records (outdoor_data_mfcc.py:62-101)."""
    from acoustic_image_generation_b200 import loader, tfrecord
    paths, all_img, all_audio, all_cls = [], [], [], []
    for r, frames in enumerate((12, 7, 20, 12, 3)):
        images = synth.smooth_images(frames, 300 + r)
        audio = synth.audio_rows(frames, 400 + r, np.int32)
        blob = tfrecord.encode_sequence_example(
            {'classes': r, 'location': 2, 'audio_image/height': 36, 'audio_image/width': 48, 'audio_image/depth': 12,
             'audio_data/mics': 1, 'audio_data/samples': 1024},
            {'audio/image': [f.tobytes() for f in images], 'audio/data': [a.tobytes() for a in audio]})
        paths.append(tfrecord.write_sequence_examples(str(tmp_path / ('Data_%03d.tfrecord' % r)), [blob]))
        all_img.append(images[:, ::-1, ::-1, :]); all_audio.append(audio); all_cls += [r] * frames
    want_img = oracle.normalize_acoustic_images(np.ascontiguousarray(np.concatenate(all_img, 0)))
    audio = np.concatenate(all_audio, 0)
    raw_mfcc = oracle.build_spectrograms(audio)
    want_mfcc = oracle.normalize_mfcc(raw_mfcc)
    want_filtered = oracle.butter_lowpass_filter(audio)
    got = list(loader.AcousticBatches(paths, batch_frames=16, low_pass=True))
    assert [len(b['classes']) for b in got] == [16, 16, 16, 6]
    assert np.concatenate([b['classes'] for b in got]).tolist() == all_cls
    cat = lambda k: torch.cat([b[k] for b in got], 0).cpu().numpy()
    assert all(b['acoustic'].is_cuda and b['mfccmap'].is_cuda for b in got)
    assert np.array_equal(cat('acoustic'), want_img)
    scale = 1.0 / (raw_mfcc.max(1) - raw_mfcc.min(1)).min()                   # min-max normalisation stretches the 1e-4
    assert np.abs(cat('mfcc') - want_mfcc).max() <= 5e-4 * max(scale, 1.0) + 1e-6
    assert np.array_equal(cat('mfccmap'), oracle.tile_mfcc(cat('mfcc')))
    assert np.array_equal(cat('filtered'), want_filtered)
    short = list(loader.AcousticBatches(paths, batch_frames=16, drop_last=True, tile=False))
    assert len(short) == 3 and 'mfccmap' not in short[0]


def test_integration_md_binding_sample_runs_verbatim():
    """The ctypes stub INTEGRATION.md shows a reference maintainer (option B) is executed as printed."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, 'INTEGRATION.md')).read()
    code = re.search(r'## B\. Bind the C ABI directly.*?```python\n(.*?)```', text, re.S).group(1)
    scope = {}
    cwd = os.getcwd()
    os.chdir(root)                                    # the sample loads the library by its repo-relative path
    try:
        exec(compile(code, 'INTEGRATION.md', 'exec'), scope)
        bank, dct, lifter, mfnorm = tables.reference_tables()
        beam = synth.power_frames(1, 0, 'chi2').reshape(-1, 512)[:100]
        got = scope['get_feats'](512, beam, 12, dct, mfnorm, lifter, bank)
    finally:
        os.chdir(cwd)
    assert np.abs(got - oracle.get_feats(512, beam, 12, dct, mfnorm, lifter, bank)).max() <= MFCC_TOL


def test_plain_c_program_against_the_abi(tmp_path):
    """examples/c_abi_smoke.c: the ABI is usable from C with nothing but the header and the shared object."""
    import os
    import subprocess
    from acoustic_image_generation_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / 'c_abi_smoke')
    subprocess.run(['gcc', '-O2', '-I', os.path.join(root, 'include'), os.path.join(root, 'examples', 'c_abi_smoke.c'), '-o', exe,
                    '-L', _lib.CSRC, '-laig', '-lm', '-Wl,-rpath,' + _lib.CSRC], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    print(out)
    assert 'tables match' in out and 'energy[0]' in out and 'auc =' in out


def test_batches_beyond_the_per_launch_limit_are_split(path):
    """TMA tile coordinates are int32, so more than 2^31 spectra (1.2 M frames) go out as several launches; the same
    split forced at 3000 spectra / 2 frames must not change a bit (with and without the 180-degree flip)."""
    power = synth.power_frames(7, 44, 'chi2')
    ref_rows = path.mfcc_rows(power.reshape(-1, 512)[:7000])
    ref_img = path.mfcc_image(power, flip=True)
    ref_chain = path.mfcc_energy(power, flip=True, normalize_first=True)
    before = path.launch_count
    path.set_option('launch_row_limit', 3000)
    try:
        import torch
        dev_rows = path.mfcc_rows(torch.from_numpy(power.reshape(-1, 512)[:7000]).cuda())
        assert path.launch_count - before == 3                      # 7000 resident rows in launches of 2816 (11 x 256)
        assert np.array_equal(dev_rows.cpu().numpy(), ref_rows)
        rows = path.mfcc_rows(power.reshape(-1, 512)[:7000])        # host rows: also cut into H2D chunks
        img = path.mfcc_image(power, flip=True)
        path.set_option('launch_row_limit', 2 * 1728)
        chain = path.mfcc_energy(power, flip=True, normalize_first=True)
    finally:
        path.set_option('launch_row_limit', (1 << 31) - 1024)
    assert np.array_equal(rows, ref_rows) and np.array_equal(img, ref_img)
    for a, b in zip(chain, ref_chain):
        assert np.array_equal(a, b)


def test_energy_heatmap_combined_call(path, golden):
    imgs = synth.smooth_images(3, 4)
    energy, mask, heat = path.energy_heatmap(imgs)
    e2, m2 = path.energy(imgs)
    assert np.array_equal(energy, e2) and np.array_equal(mask, m2)
    assert np.array_equal(heat, path.heatmap(e2))
    assert np.abs(heat[0] - oracle.heatmap(oracle.find_logen(imgs[0].copy()))).max() <= 2e-6


def test_render_heatmaps_driver(path):
    from acoustic_image_generation_b200 import evaluate
    imgs = synth.smooth_images(2, 6)
    frames = np.random.default_rng(3).integers(0, 256, (2, 224, 298, 3), dtype=np.uint8)
    rgb = evaluate.render_heatmaps(path, imgs, frames)
    assert rgb.shape == (2, 224, 298, 3) and rgb.dtype == np.uint8
    heat = path.heatmap(path.energy(imgs)[0])
    assert np.array_equal(rgb[1], oracle.overlay(heat[1], frames[1], tables.jet_lut()))


def test_generic_tables_with_flip_and_zero_thresholds(path):
    bank2 = aig.createfilters(128, 10, 0, 3000, 8000)
    dct2, lifter2, mfnorm2 = tables.mfcc_constants(10, 6, 22)
    p2 = aig.AcousticPath(0, tables_=(bank2, dct2, lifter2, mfnorm2))
    rng = np.random.default_rng(9)
    beam = rng.standard_normal((35, 128), dtype=np.float32) ** 2          # 5 "frames" of 7 rows
    plain = p2.mfcc_rows(beam)
    flipped = p2.mfcc_rows(beam, flip180=True, frame_pixels=7)
    assert np.array_equal(flipped.reshape(5, 7, 6), plain.reshape(5, 7, 6)[:, ::-1, :])
    want = oracle.get_feats(128, beam, 6, dct2, mfnorm2, lifter2, bank2)
    assert np.abs(plain - np.float32(want)).max() <= 1e-5
    p2.close()
    # a sweep with no thresholds still returns the per-frame counts
    m = (rng.random((3, 36, 48)) > 0.5).astype(np.uint8)
    inter, union, pos, num = path.iou_sweep(m, m[::-1].copy(), [])
    assert pos.shape == (0,) and num == 3 and inter.tolist() == [oracle.iou_pair(a, b)[0] for a, b in zip(m, m[::-1])]


def test_dlpack_and_cuda_array_interface_inputs(path, torch):
    class DlpackOnly:                      # what a CuPy / JAX / TF array looks like to us
        def __init__(self, t):
            self._t = t

        def __dlpack__(self, *args, **kwargs):
            return self._t.__dlpack__(*args, **kwargs)

        def __dlpack_device__(self):
            return self._t.__dlpack_device__()

    class CaiOnly:
        def __init__(self, t):
            self._t = t
            self.__cuda_array_interface__ = t.__cuda_array_interface__

    imgs = torch.from_numpy(synth.sigmoid_images(2, 12)).cuda()
    ref_e, ref_m = path.energy(imgs)
    for wrapper in (DlpackOnly, CaiOnly):
        e, m = path.energy(wrapper(imgs))
        assert e.is_cuda and torch.equal(e, ref_e) and torch.equal(m, ref_m)


def test_tfrecord_to_metric_files_example(tmp_path):
    """examples/evaluate_tfrecords.py: TFRecords -> reader -> normalise -> energy masks -> IoU sweep -> metric files,
    checked against the oracle chain on the same records."""
    import importlib.util
    import os
    from acoustic_image_generation_b200 import metrics_io, tfrecord
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('evaluate_tfrecords', os.path.join(root, 'examples', 'evaluate_tfrecords.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    res, out_dir = mod.main(str(tmp_path))
    assert res['num'] == 72
    rng = np.random.default_rng(0)
    scores = []
    for r in range(6):
        with tfrecord.RecordFile(os.path.join(out_dir, 'Data_%03d.tfrecord' % (r + 1))) as rec:
            imgs = tfrecord.parse_acoustic_example(rec, 0)['audio_images']
        assert np.array_equal(imgs, oracle.flip180(synth.smooth_images(12, 100 + r)))
        data = oracle.normalize_acoustic_images(imgs)
        recon = np.clip(data * np.float32(0.8) + np.float32(0.2) * rng.random(data.shape, dtype=np.float32), 0, 1)
        ea, _ = oracle.energy_stage(data, normalize_first=False)
        eb, _ = oracle.energy_stage(recon, normalize_first=False)
        scores += [oracle.iou_pair(oracle.mean_mask(x), oracle.mean_mask(y))[2] for x, y in zip(ea, eb)]
    pos, num = oracle.success_counts(scores, REF_THR)
    assert np.array_equal(res['pos'], pos) and num == 72
    assert metrics_io.read_accuracy_file(out_dir, 0.5) == float('{:6f}'.format(pos[5] / 72))


def test_energy_stage_arithmetic_shortcuts_selftest(path):
    """The energy stage divides by the lifter with a reciprocal + one exact FMA correction (Markstein) and evaluates exp
    with a 1024-entry table on arguments in table-step units; both are checked on the device: the division against IEEE
    division for every float32 input and all twelve lifters, the exp against CUDA's exp() on 1.3e8 points.  The exp
    reference is exp(u_hi) * (1 + u_lo) with CUDA's exp (documented <= 1 ulp), so 2 ulp between the two is the bound a
    <= 1 ulp table exp can show; tests/test_energy_tables_cpu.py measures the table exp itself in exact arithmetic."""
    bad, bad_after_store, _, _ = path.selftest(0)
    print('division: %d of %d (input, lifter) pairs differ from IEEE division; %d after the float32 store' % (bad, 12 << 32, bad_after_store))
    assert bad == 0 and bad_after_store == 0
    differ, max_ulps, count, _ = path.selftest(1)
    print('exp: %d of %d points differ from CUDA exp(), max %d ulp' % (differ, count, max_ulps))
    assert count == 2 << 26 and max_ulps <= 2
