"""BASELINE.json configs[2] and configs[3] at their stated sizes, through the C ABI on a B200.

The oracle runs on a random subset (it needs ~1 ms per frame); the full-size results are held to
size-independent properties: the sweep counts equal the counts recomputed from the per-frame integer (I, U),
the success curve is non-increasing, shards merge to the global result, AUC of merged counts == global AUC.
"""
import numpy as np
import pytest

import acoustic_image_generation_b200 as aig
from acoustic_image_generation_b200 import sharding, synth
from oracle import acoustic_oracle as oracle

pytestmark = pytest.mark.gpu
REF_THR = list(oracle.REFERENCE_THRESHOLDS)


@pytest.fixture(scope='module')
def path():
    p = aig.AcousticPath(0)
    yield p
    p.close()


def counts_from_ratios(num, den, thr):
    with np.errstate(invalid='ignore', divide='ignore'):
        score = num.astype(np.float64) / den.astype(np.float64)
    return oracle.success_counts(score, thr)[0]


@pytest.mark.parametrize('shape', [(224, 224), (224, 298)])
def test_config2_flickr_5k_frames_heatmap_and_ciou_sweep(path, shape):
    """configs[2]: energy heat map up-sampled to 224x224 (BASELINE) / 224x298 (reference) + consensus-IoU sweep over
    101 thresholds (and the reference's 11) on FlickrSoundNet-shaped boxes, 5000 frames."""
    n = 5000
    pred = synth.smooth_images(n, 60)
    boxes = synth.flickr_boxes(n, 61, shape[0], shape[1])
    energy, masks = path.energy(pred)
    heat = path.heatmap(energy, *shape)
    assert heat.shape == (n,) + shape and heat.dtype == np.float32
    assert np.all(heat.reshape(n, -1).min(1) == 0.0) and np.all(np.abs(heat.reshape(n, -1).max(1) - 1.0) <= 1e-6)
    thr101 = np.linspace(0, 1, 101)
    i2, u2, pos101, num = path.ciou_sweep(masks, *boxes, thr101, out_hw=shape)
    _, _, pos11, num11 = path.ciou_sweep(masks, *boxes, REF_THR, out_hw=shape)
    assert num == n and num11 == n
    assert np.array_equal(pos101, counts_from_ratios(i2, u2, thr101))
    assert np.array_equal(pos11, counts_from_ratios(i2, u2, REF_THR))
    assert np.all(np.diff(pos101) <= 0) and pos101[-1] == 0
    assert np.all(i2 <= u2) and np.all(u2 <= 2 * shape[0] * shape[1])
    # oracle on a random subset
    rng = np.random.default_rng(0)
    for h in rng.choice(n, 48, replace=False):
        want_e = oracle.find_logen(pred[h].copy())
        assert np.abs(energy[h] - want_e).max() <= 1e-13 * np.abs(want_e).max()
        want_m = oracle.mean_mask(want_e)
        assert np.array_equal(masks[h], want_m)
        assert np.abs(heat[h] - oracle.heatmap(want_e, *shape)).max() <= 1e-4
        gt = oracle.boxes_to_consensus(boxes[0][h], boxes[1][h], boxes[2][h], boxes[3][h], *shape)
        wi, wu, _ = oracle.consensus_iou(gt, oracle.resize_mask(want_m, *shape))
        assert (int(i2[h]), int(u2[h])) == (wi, wu)
    auc101 = aig.auc(thr101, aig.success_rates(pos101, num))
    assert abs(auc101 - oracle.auc(thr101, oracle.success_rates(pos101, num))) <= 1e-12
    assert 0.0 <= auc101 <= 1.0


@pytest.mark.parametrize('frames_per_clip,n_clips', [(120, 16), (300, 8)])
def test_config3_clip_stream_sharded_by_clip(path, frames_per_clip, n_clips):
    """configs[3]: VGGSound-shaped stream of 10 s clips (120 frames at the reference's 12 fps, 300 at BASELINE's
    30 fps): MFCC + energy + per-clip and global AUC; clips sharded over 2 / 4 / 8 ranks merge to the global counts."""
    import torch
    n = frames_per_clip * n_clips
    base = torch.from_numpy(synth.power_frames(24, 70, 'chi2')).cuda()
    idx = torch.from_numpy(np.random.default_rng(1).integers(0, 24, n)).cuda()
    power_a = base[idx]                                   # "real" stream
    power_b = base[torch.roll(idx, 1)]                    # "generated" stream: neighbouring frame
    _, _, mask_a = path.mfcc_energy(power_a, flip=True, normalize_first=True)
    _, _, mask_b = path.mfcc_energy(power_b, flip=True, normalize_first=True)
    inter, union, pos, num = path.iou_sweep(mask_a, mask_b, REF_THR)
    inter, union = inter.cpu().numpy(), union.cpu().numpy()
    assert num == n and np.array_equal(pos, counts_from_ratios(inter, union, REF_THR))
    global_auc = aig.auc(REF_THR, aig.success_rates(pos, num))
    # per-clip AUC from per-clip sweeps == AUC recomputed from the per-frame (I, U)
    for c in range(0, n_clips, max(1, n_clips // 4)):
        lo, hi = c * frames_per_clip, (c + 1) * frames_per_clip
        _, _, cpos, cnum = path.iou_sweep(mask_a[lo:hi], mask_b[lo:hi], REF_THR)
        assert cnum == frames_per_clip
        assert np.array_equal(cpos, counts_from_ratios(inter[lo:hi], union[lo:hi], REF_THR))
    # ... and all clips at once from the per-clip kernel (one launch)
    _, _, clip_pos = path.iou_sweep_clips(mask_a, mask_b, frames_per_clip, REF_THR)
    assert clip_pos.shape == (n_clips, len(REF_THR)) and np.array_equal(clip_pos.sum(axis=0), pos)
    for c in range(n_clips):
        lo, hi = c * frames_per_clip, (c + 1) * frames_per_clip
        assert np.array_equal(clip_pos[c], counts_from_ratios(inter[lo:hi], union[lo:hi], REF_THR))
    clip_auc = [aig.auc(REF_THR, aig.success_rates(p, frames_per_clip)) for p in clip_pos]
    assert min(clip_auc) <= global_auc <= max(clip_auc)
    # a ragged last clip
    _, _, ragged = path.iou_sweep_clips(mask_a[:frames_per_clip + 7], mask_b[:frames_per_clip + 7], frames_per_clip, REF_THR)
    assert ragged.shape == (2, len(REF_THR)) and np.array_equal(ragged[0], clip_pos[0])
    assert np.array_equal(ragged[1], counts_from_ratios(inter[frames_per_clip:frames_per_clip + 7], union[frames_per_clip:frames_per_clip + 7], REF_THR))
    # clip-major shards merge to the global result
    for world in (2, 4, 8):
        shards = []
        for rank in range(world):
            c0, c1, f0, f1 = sharding.shard_clips(n_clips, frames_per_clip, rank, world)
            if f1 > f0:
                _, _, spos, snum = path.iou_sweep(mask_a[f0:f1], mask_b[f0:f1], REF_THR)
            else:
                spos, snum = np.zeros(len(REF_THR), np.int64), 0
            shards.append(np.concatenate([spos, [snum]]))
        merged = sharding.merge_counts(shards)
        assert np.array_equal(merged[:-1], pos) and merged[-1] == n
        assert aig.auc(REF_THR, aig.success_rates(merged[:-1], merged[-1])) == global_auc
    # oracle on a few frames of the stream
    want = oracle.mfcc_image(base.cpu().numpy(), flip=True)
    _, want_masks = oracle.energy_stage(want, normalize_first=True)
    host_idx = idx.cpu().numpy()
    flips = 0
    for h in range(0, n, max(1, n // 40)):
        differing = np.argwhere(mask_a[h].cpu().numpy() != want_masks[host_idx[h]])
        flips += len(differing)
        for y, x in differing:                  # north star: every boundary-pixel disagreement is reported
            print('  boundary pixel: stream frame %d (base frame %d) pixel (%d, %d)' % (h, host_idx[h], y, x))
    print('clip stream: boundary-pixel disagreements on sampled frames: %d' % flips)
    assert flips <= 8


def test_soak_2048_distinct_frames_invariants_and_oracle_sample(path):
    """2048 distinct random frames (7.2 GB) through the fused pass: global invariants recomputed on the host from the
    GPU's own float64 energies for every frame, the oracle on a 48-frame sample, and the degenerate frames the reference
    turns into NaN."""
    import torch
    n = 2048
    gen = torch.Generator(device='cuda')
    gen.manual_seed(99)
    power = torch.empty((n, 36, 48, 512), device='cuda', dtype=torch.float32)
    power.normal_(generator=gen).square_()
    power[5] = float('nan')                     # every spectrum non-finite -> MFCC all 0 -> constant frame -> 0/0
    power[6] = 0.0                              # all-floor frame: MFCC ~ 1e-6 rounding noise, still normalisable
    mfcc, energy, mask, mean = path.mfcc_energy(power, flip=True, normalize_first=True, want_mean=True)
    mfcc_h, energy_h, mask_h, mean_h = (t.cpu().numpy() for t in (mfcc, energy, mask, mean))
    assert np.all(mfcc_h[5] == 0.0) and np.isnan(energy_h[5]).all() and np.isnan(mean_h[5]) and mask_h[5].sum() == 0
    ok = np.ones(n, bool); ok[5] = False
    assert np.isfinite(mfcc_h).all() and np.isfinite(energy_h[ok]).all()
    # mean and mask are functions of the energies alone: recompute them with NumPy for all frames
    assert np.array_equal(mean_h[ok], energy_h[ok].reshape(-1, 1728).mean(axis=1))
    assert np.array_equal(mask_h[ok], (energy_h[ok] > mean_h[ok][:, None, None]).astype(np.uint8))
    assert 0.2 < mask_h[ok].mean() < 0.8
    # energies are functions of the MFCC image alone: replay stage 2 of the oracle on the GPU's MFCC for a sample
    rng = np.random.default_rng(7)
    sample = [5, 6] + rng.choice(n, 46, replace=False).tolist()
    host_power = power[sample].cpu().numpy()
    with np.errstate(invalid='ignore'):
        want_mfcc = oracle.mfcc_image(host_power, flip=True)
        err = np.abs(mfcc_h[sample] - want_mfcc).max()
        assert err <= 1e-4
        e_replay, m_replay = oracle.energy_stage(mfcc_h[sample], normalize_first=True)
    fin = np.isfinite(e_replay).all(axis=(1, 2))
    assert not fin[0] and fin[1:].all()
    assert np.abs(energy_h[sample][fin] - e_replay[fin]).max() <= 1e-13 * np.abs(e_replay[fin]).max()
    assert np.array_equal(mask_h[sample][fin], m_replay[fin])
    # and end to end against the oracle's own MFCC: count boundary flips
    with np.errstate(invalid='ignore'):
        e_oracle, m_oracle = oracle.energy_stage(want_mfcc, normalize_first=True)
    per_frame = (mask_h[sample] != m_oracle).reshape(len(sample), -1).sum(axis=1)
    # north star: every boundary-pixel disagreement is reported - frame, pixel, both energies and both frame means
    for k, y, x in np.argwhere(mask_h[sample][2:] != m_oracle[2:]):
        f = sample[2 + k]
        print('  boundary pixel: frame %d (%d, %d)  gpu energy %.17g mean %.17g | oracle energy %.17g mean %.17g'
              % (f, y, x, energy_h[f, y, x], mean_h[f], e_oracle[2 + k, y, x], e_oracle[2 + k].mean()))
    # sample[1] is the all-floor frame: its MFCC is mathematically 0, i.e. pure rounding noise (1e-15 in the reference's
    # float64, 1e-6 here, both far inside the 1e-4 tolerance) that the min-max normalisation then stretches to [0, 1] -
    # the reference's own mask for such a frame is noise, so it is excluded from the boundary-pixel count
    flips = int(per_frame[2:].sum())
    print('soak: mfcc max abs err %.2e; boundary flips vs the all-oracle chain on %d frames: %d (all-floor frame: %d)'
          % (err, len(sample) - 2, flips, int(per_frame[1])))
    assert flips <= 10


def test_mfcc_error_statistics_over_many_spectra(path):
    """Error distribution of the float32 MFCC kernel against the float64 reference arithmetic over 3.3e5 spectra per input
    family (the golden fixtures cover 1-2 frames): max, 99.99th percentile and RMS of the absolute error, all far inside
    the 1e-4 tolerance."""
    bank, dct, lifter, mfnorm = oracle.reference_tables()
    rng = np.random.default_rng(123)
    families = {
        'chi2': synth.power_frames(64, 301, 'chi2').reshape(-1, 512),
        'lognormal sigma=3': synth.power_frames(64, 302, 'lognormal').reshape(-1, 512),
        'near the 0.001 floor': (rng.random((64 * 1728, 512), dtype=np.float32) * np.float32(4e-4)),
        'one dominant bin': (rng.random((64 * 1728, 512), dtype=np.float32) * np.float32(1e-3)
                             + (rng.random((64 * 1728, 512)) > 0.998).astype(np.float32) * np.float32(1e6)),
        'audio-scale 1e9': synth.power_frames(64, 303, 'chi2').reshape(-1, 512) * np.float32(1e9),
    }
    for name, beam in families.items():
        want = oracle.get_feats(512, beam, 12, dct, mfnorm, lifter, bank)
        got = path.mfcc_rows(beam).astype(np.float64)
        err = np.abs(got - np.float32(want).astype(np.float64)).ravel()
        print('%-22s n=%d  max %.2e  p99.99 %.2e  rms %.2e  (max |mfcc| %.1f)'
              % (name, err.size, err.max(), np.quantile(err, 0.9999), np.sqrt(np.mean(err ** 2)), np.abs(want).max()))
        assert err.max() <= 1e-4


def test_full_size_properties_8192_frames(path):
    """configs[1] at the size the bench runs (8192 resident frames, 29 GB): properties that need no oracle.
    Determinism of the persistent fused kernel, flip equivariance, batch-split invariance, frame independence (bitwise),
    and invariance of the MFCC to a global gain (the DCT rows m >= 1 annihilate constants; tolerance 1e-4)."""
    import torch
    n = 8192
    gen = torch.Generator(device='cuda')
    gen.manual_seed(2024)
    power = torch.empty((n, 36, 48, 512), device='cuda', dtype=torch.float32)
    for f0 in range(0, n, 512):
        power[f0:f0 + 512].normal_(generator=gen).square_()
    mfcc, energy, mask = path.mfcc_energy(power, flip=True, normalize_first=True)
    # (a) determinism
    mfcc2, energy2, mask2 = path.mfcc_energy(power, flip=True, normalize_first=True)
    assert torch.equal(mfcc, mfcc2) and torch.equal(mask, mask2) and torch.equal(energy.view(torch.int64), energy2.view(torch.int64))
    del mfcc2, energy2, mask2
    # (b) flip equivariance: the 180-degree rotation is pure store addressing
    plain = path.mfcc_image(power, flip=False)
    assert torch.equal(plain.flip(1, 2), mfcc)
    del plain
    # (c) batch-split invariance and (d) frame independence
    lo = path.mfcc_energy(power[:3000], flip=True, normalize_first=True)
    hi = path.mfcc_energy(power[3000:], flip=True, normalize_first=True)
    assert torch.equal(torch.cat([lo[0], hi[0]]), mfcc) and torch.equal(torch.cat([lo[2], hi[2]]), mask)
    assert torch.equal(torch.cat([lo[1], hi[1]]).view(torch.int64), energy.view(torch.int64))
    del lo, hi
    for f in (0, 147, 148, 4095, 8191):
        one = path.mfcc_energy(power[f:f + 1], flip=True, normalize_first=True)
        assert torch.equal(one[0][0], mfcc[f]) and torch.equal(one[2][0], mask[f])
        assert torch.equal(one[1][0].view(torch.int64), energy[f].view(torch.int64))
    # (e) gain invariance: x4 is exact in float32 and every mel band stays far above the 0.001 floor
    power.mul_(4.0)
    gained = path.mfcc_image(power, flip=True)
    err = float((gained - mfcc).abs().max())
    print('full size: MFCC change under a global gain of 4: %.2e' % err)
    assert err <= 1e-4
    # (f) the masks are balanced and the IoU sweep is symmetric in its two streams
    assert 0.3 < float(mask.float().mean()) < 0.7
    half = n // 2
    a = path.iou_sweep(mask[:half], mask[half:])
    b = path.iou_sweep(mask[half:], mask[:half])
    assert all(torch.equal(torch.as_tensor(x), torch.as_tensor(y)) for x, y in zip(a, b))
