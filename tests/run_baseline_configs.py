"""Run the five BASELINE.json configs on one B200 and print one JSON line each (frames/s + what was verified).

    python tests/run_baseline_configs.py [--quick]      (lives under tests/ because it uses the oracle as its checker)

C1  16 frames vs the oracle (the reference's CPU-runnable case)
C2  1 M frames MFCC + energy: a resident ring of 8192 frames cycled 123 times
C3  5 k FlickrSoundNet-shaped frames: energy, heat map 224x224, consensus-IoU sweep over 101 thresholds
C4  VGGSound-shaped clip stream (120 / 300 frames per clip): MFCC + energy + per-clip and global AUC
C5  10 M frames end to end (MFCC + energy + IoU + AUC): the same ring cycled 1221 times with device-resident counters
    (on N GPUs each rank runs 1/N of the cycles and the counters are all-reduced: bench.py --gpus N)
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import acoustic_image_generation_b200 as aig  # noqa: E402
from acoustic_image_generation_b200 import synth  # noqa: E402
from oracle import acoustic_oracle as oracle  # noqa: E402  (checker only)

THR11 = list(aig.REFERENCE_THRESHOLDS)


def timed(stream, fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    out = fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1), out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--quick', action='store_true', help='1/10 of the cycles for C2 and C5')
    args = ap.parse_args()
    dev = torch.device('cuda', 0)
    stream = torch.cuda.current_stream()
    path = aig.AcousticPath(0, stream=stream.cuda_stream)

    # ---- C1 -------------------------------------------------------------------------------------------------
    power = synth.power_frames(16, 0, 'chi2')
    t0 = time.perf_counter()
    mfcc, energy, mask = path.mfcc_energy(power, flip=True, normalize_first=True)
    dt = time.perf_counter() - t0
    t0 = time.perf_counter()
    for _ in range(20):                                                     # steady state, ordinary NumPy in / out
        path.mfcc_energy(power, flip=True, normalize_first=True)
    steady = (time.perf_counter() - t0) / 20
    t0 = time.perf_counter()
    for _ in range(3):
        oracle.energy_stage(oracle.mfcc_image(power, flip=True), normalize_first=True)
    cpu = (time.perf_counter() - t0) / 3
    want = oracle.mfcc_image(power, flip=True)
    _, want_mask = oracle.energy_stage(want, normalize_first=True)
    print(json.dumps({'config': 'C1: 16 frames 36x48x512 -> 12-ch MFCC (+energy, mask), host in / host out',
                      'frames_per_s_first_call': 16 / dt, 'frames_per_s_steady': 16 / steady, 'ms_per_call': 1e3 * steady,
                      'numpy_oracle_frames_per_s_same_host': 16 / cpu, 'mfcc_max_abs_err': float(np.abs(mfcc - want).max()),
                      'mask_pixels_differing': int((mask != want_mask).sum())}), flush=True)

    # ---- resident ring for C2 / C5 ---------------------------------------------------------------------------
    n = 8192
    gen = torch.Generator(device=dev); gen.manual_seed(1)
    ring = torch.empty((n, 36, 48, 512), device=dev, dtype=torch.float32)
    for f0 in range(0, n, 256):
        ring[f0:f0 + 256].normal_(generator=gen).square_()
    out = (torch.empty((n, 36, 48, 12), device=dev, dtype=torch.float32), torch.empty((n, 36, 48), device=dev, dtype=torch.float64),
           torch.empty((n, 36, 48), device=dev, dtype=torch.uint8))
    for _ in range(3):
        path.mfcc_energy(ring, flip=True, normalize_first=True, out=out)
    cycles = 123 // (10 if args.quick else 1)
    ms, _ = timed(stream, lambda: [path.mfcc_energy(ring, flip=True, normalize_first=True, out=out) for _ in range(cycles)])
    print(json.dumps({'config': 'C2: MFCC + energy for %d frames (ring of %d resident frames = 29 GB cycled %d x)' % (n * cycles, n, cycles),
                      'frames_per_s': n * cycles / ms * 1e3, 'algorithmic_GBps': n * cycles * 3628800 / ms / 1e6,
                      'seconds': ms / 1e3}), flush=True)

    # ---- C5 --------------------------------------------------------------------------------------------------
    thr = torch.tensor(THR11, device=dev, dtype=torch.float64)
    counts = torch.zeros(12, device=dev, dtype=torch.int64)
    half = n // 2
    cycles5 = 1221 // (10 if args.quick else 1)

    def full_chain():
        for _ in range(cycles5):
            path.mfcc_energy(ring, flip=True, normalize_first=True, out=out)
            path.iou_sweep(out[2][:half], out[2][half:], thr, pos=counts[:-1], num=counts[-1:])
        path.allreduce_counts(counts)
    ms, _ = timed(stream, full_chain)
    host = counts.cpu().numpy()
    auc = aig.auc(THR11, aig.success_rates(host[:-1], int(host[-1])))
    print(json.dumps({'config': 'C5: %d frames MFCC + energy + IoU pairs + AUC, device-resident counters' % (n * cycles5),
                      'frames_per_s': n * cycles5 / ms * 1e3, 'seconds': ms / 1e3, 'pairs_scored': int(host[-1]), 'auc': auc,
                      'pos': host[:-1].tolist()}), flush=True)
    del ring, out

    # ---- C3 --------------------------------------------------------------------------------------------------
    m = 5000
    pred = torch.from_numpy(synth.smooth_images(m, 60)).to(dev)
    boxes = [torch.from_numpy(b).to(dev) for b in synth.flickr_boxes(m, 61, 224, 224)]
    thr101 = torch.linspace(0, 1, 101, device=dev, dtype=torch.float64)
    c101 = torch.zeros(102, device=dev, dtype=torch.int64)

    def flickr():
        energy, mask, heat = path.energy_heatmap(pred, False, 224, 224)
        i2, u2, _, _ = path.ciou_sweep(mask, *boxes, thr101, out_hw=(224, 224), pos=c101[:-1], num=c101[-1:])
        return energy, mask, heat, i2, u2
    flickr(); c101.zero_()
    ms, (energy, mask, heat, i2, u2) = timed(stream, flickr)
    h101 = c101.cpu().numpy()
    sample = np.random.default_rng(0).choice(m, 32, replace=False)
    pred_h, boxes_h = pred.cpu().numpy(), [b.cpu().numpy() for b in boxes]
    bad = 0
    for hh in sample:
        e = oracle.find_logen(pred_h[hh].copy())
        gt = oracle.boxes_to_consensus(boxes_h[0][hh], boxes_h[1][hh], boxes_h[2][hh], boxes_h[3][hh], 224, 224)
        wi, wu, _ = oracle.consensus_iou(gt, oracle.resize_mask(oracle.mean_mask(e), 224, 224))
        bad += (int(i2[hh]), int(u2[hh])) != (wi, wu)
    print(json.dumps({'config': 'C3: 5000 frames energy + heat map 224x224 + consensus-IoU sweep, 101 thresholds',
                      'frames_per_s': m / ms * 1e3, 'ms': ms, 'auc101': aig.auc(np.linspace(0, 1, 101), aig.success_rates(h101[:-1], int(h101[-1]))),
                      'oracle_sample_frames': 32, 'oracle_sample_IU_mismatches': int(bad)}), flush=True)

    # ---- C4 --------------------------------------------------------------------------------------------------
    for fpc, clips in ((120, 32), (300, 16)):
        nf = fpc * clips
        base = torch.from_numpy(synth.power_frames(24, 70, 'chi2')).to(dev)
        idx = torch.from_numpy(np.random.default_rng(1).integers(0, 24, nf)).to(dev)
        a, b = base[idx], base[torch.roll(idx, 1)]
        path.mfcc_energy(a[:64], flip=True, normalize_first=True)

        def clip_stream():
            _, _, ma = path.mfcc_energy(a, flip=True, normalize_first=True)
            _, _, mb = path.mfcc_energy(b, flip=True, normalize_first=True)
            _, _, clip_pos = path.iou_sweep_clips(ma, mb, fpc, THR11)              # every clip's counts in one launch
            clip_pos = clip_pos if isinstance(clip_pos, np.ndarray) else clip_pos.cpu().numpy()
            per_clip = [aig.auc(THR11, aig.success_rates(p, fpc)) for p in clip_pos]
            return per_clip, aig.auc(THR11, aig.success_rates(clip_pos.sum(axis=0), nf))
        ms, (per_clip, glob) = timed(stream, clip_stream)
        print(json.dumps({'config': 'C4: %d clips x %d frames (two streams), MFCC + energy + per-clip and global AUC' % (clips, fpc),
                          'frames_per_s': 2 * nf / ms * 1e3, 'ms': ms, 'global_auc': glob,
                          'per_clip_auc_min_max': [min(per_clip), max(per_clip)]}), flush=True)


if __name__ == '__main__':
    main()
