"""AcousticPathGroup: one process, one host call, frames sharded over per-device handles and threads.  Needs a B200; the
two-device cases need two (they are skipped on the driver's one-GPU test box, where the one-device group still runs the
same code path: threads, sharding, device-resident counters, reduction)."""
import numpy as np
import pytest

import acoustic_image_generation_b200 as aig
from acoustic_image_generation_b200 import synth
from acoustic_image_generation_b200.group import AcousticPathGroup
from oracle import acoustic_oracle as oracle

pytestmark = pytest.mark.gpu
REF_THR = list(oracle.REFERENCE_THRESHOLDS)


def _device_count():
    import torch
    return torch.cuda.device_count()


def _reference_results(power, real, recon, boxes):
    p = aig.AcousticPath(0)
    try:
        chain = p.mfcc_energy(power, flip=True, normalize_first=True)
        inter, union, pos, num = p.acivw_batch(real, recon, REF_THR)
        _, mask = p.energy(recon)
        i2, u2, cpos, cnum = p.ciou_sweep(mask, *boxes, REF_THR)
        _, _, heat = p.energy_heatmap(recon, want_energy=False, want_mask=False)
    finally:
        p.close()
    return chain, (inter, union, pos, num), (i2, u2, cpos, cnum), heat


@pytest.mark.parametrize('n_devices', [1, 2])
def test_group_results_equal_one_gpu_bitwise(n_devices):
    if _device_count() < n_devices:
        pytest.skip('needs %d GPUs' % n_devices)
    power = synth.power_frames(7, 90, 'chi2')
    real, recon = synth.smooth_images(37, 91), synth.smooth_images(37, 92)
    recon[:10] = real[:10] * np.float32(0.8) + recon[:10] * np.float32(0.2)
    boxes = synth.flickr_boxes(37, 93)
    chain, acivw, flickr, heat = _reference_results(power, real, recon, boxes)
    with AcousticPathGroup(list(range(n_devices)), thresholds=REF_THR) as group:
        assert len(group) == n_devices and (group.nccl or n_devices == 1)
        got = group.mfcc_energy(power, flip=True, normalize_first=True)
        for a, b in zip(got, chain):
            assert np.array_equal(a, b)
        dyn = group.mfcc_energy(power, flip=True, normalize_first=True, chunk_frames=2)     # pieces taken dynamically
        for a, b in zip(dyn, chain):
            assert np.array_equal(a, b)
        inter, union = group.add_acivw_batch(real, recon)
        assert np.array_equal(inter, acivw[0]) and np.array_equal(union, acivw[1])
        res = group.finish()
        assert np.array_equal(res['pos'], acivw[2]) and res['num'] == acivw[3]
        # a second batch accumulates on top of the reduced counts exactly once
        group.add_acivw_batch(real[:5], recon[:5])
        res2 = group.finish()
        assert res2['num'] == acivw[3] + 5
        group.reset_counts()
        i2, u2 = group.add_flickr_batch(recon, *boxes)
        assert np.array_equal(i2, flickr[0]) and np.array_equal(u2, flickr[1])
        res = group.finish()
        assert np.array_equal(res['pos'], flickr[2]) and res['num'] == flickr[3]
        assert np.array_equal(group.energy_heatmap(recon), heat)


def test_group_host_sum_equals_nccl_sum():
    if _device_count() < 2:
        pytest.skip('needs 2 GPUs')
    real, recon = synth.smooth_images(20, 94), synth.smooth_images(20, 95)
    totals = []
    for use_nccl in (True, False):
        with AcousticPathGroup([0, 1], thresholds=REF_THR, use_nccl=use_nccl) as group:
            assert group.nccl == use_nccl
            group.add_acivw_batch(real, recon)
            totals.append(group.reduce_counts())
    assert np.array_equal(totals[0], totals[1])


def test_group_rejects_bad_device_lists():
    with pytest.raises(ValueError):
        AcousticPathGroup([0, 0])
    with pytest.raises(ValueError):
        AcousticPathGroup([])
