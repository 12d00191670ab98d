"""Per-stage timings of every kernel behind the C ABI on a B200, with the roofline that bounds each
(resident inputs, CUDA events on the handle's stream, median of --iters).

    python tools/bench_stages.py [--frames 4096] [--iters 7]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import acoustic_image_generation_b200 as aig  # noqa: E402
from acoustic_image_generation_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--frames', type=int, default=4096)
    ap.add_argument('--iters', type=int, default=7)
    args = ap.parse_args()
    peak = 6545.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
    except Exception:
        pass
    dev = torch.device('cuda', 0)
    stream = torch.cuda.current_stream()
    path = aig.AcousticPath(0, stream=stream.cuda_stream)
    n = args.frames

    def timed(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        times = []
        for _ in range(args.iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        return sorted(times)[len(times) // 2]

    rows = []

    def report(name, units, unit_name, ms, bytes_moved, note=''):
        gbs = bytes_moved / ms / 1e6
        rows.append((name, units, unit_name, ms, gbs))
        print('%-34s %9d %-8s %9.3f ms %12.0f %s/s %8.0f GB/s (%.2f of measured HBM peak) %s'
              % (name, units, unit_name, ms, units / ms * 1e3, unit_name, gbs, gbs / peak, note), flush=True)

    power = torch.randn((n, 36, 48, 512), device=dev, dtype=torch.float32).square_()
    mfcc = torch.empty((n, 36, 48, 12), device=dev, dtype=torch.float32)
    energy = torch.empty((n, 36, 48), device=dev, dtype=torch.float64)
    mask = torch.empty((n, 36, 48), device=dev, dtype=torch.uint8)
    ms = timed(lambda: path.mfcc_energy(power, flip=True, normalize_first=True, out=(mfcc, energy, mask)))
    report('fused MFCC+energy (aig_mfcc_energy)', n, 'frames', ms, n * 3628800)
    ms = timed(lambda: path.mfcc_rows(power, flip180=True, out=mfcc.view(-1, 12)))
    report('MFCC alone (aig_mfcc)', n, 'frames', ms, n * (3538944 + 82944))
    del power
    ms = timed(lambda: path.energy(mfcc, normalize_first=True))
    report('energy+mask alone (aig_energy)', n, 'frames', ms, n * (82944 + 13824 + 1728), 'FP64-bound')
    other = torch.rand((n, 36, 48, 12), device=dev, dtype=torch.float32)
    thr11 = torch.tensor(aig.REFERENCE_THRESHOLDS, device=dev, dtype=torch.float64)
    cnt = torch.zeros(12, device=dev, dtype=torch.int64)
    ms = timed(lambda: path.acivw_batch(mfcc, other, thr11, pos=cnt[:-1], num=cnt[-1:]))
    report('ACIVW step, one launch (aig_acivw_batch)', n, 'pairs', ms, n * 2 * 82944, 'FP64-bound: two energy maps per pair')
    for small in (1, 2, 16, 64):
        ms = timed(lambda: path.energy(mfcc[:small]))
        report('energy, cluster form, %d frames' % small, small, 'frames', ms, small * (82944 + 13824 + 1728), 'latency')
        ms = timed(lambda: path.acivw_batch(mfcc[:small], other[:small], thr11, pos=cnt[:-1], num=cnt[-1:]))
        report('ACIVW step, cluster form, %d pairs' % small, small, 'pairs', ms, small * 2 * 82944, 'latency')
    del other
    ms = timed(lambda: path.normalize_images(mfcc))
    report('min-max normalise (aig_normalize)', n, 'frames', ms, n * 2 * 82944)
    m = min(n, 2048)
    for shape in ((224, 298), (224, 224)):
        out = torch.empty((n,) + shape, device=dev, dtype=torch.float32)
        for count in sorted({m, n}):
            ms = timed(lambda: path._lib.aig_heatmap(path._h, energy.data_ptr(), count, shape[0], shape[1], out.data_ptr()))
            report('heat map %dx%d (aig_heatmap, bulk copies)' % shape, count, 'frames', ms, count * (13824 + shape[0] * shape[1] * 4),
                   'write roof 6.3 TB/s: %.2f' % (count * shape[0] * shape[1] * 4 / ms / 1e6 / 6300.0))
        path.set_option('heat_bulk_store', 0)
        ms = timed(lambda: path._lib.aig_heatmap(path._h, energy.data_ptr(), n, shape[0], shape[1], out.data_ptr()))
        path.set_option('heat_bulk_store', 1)
        report('heat map %dx%d (round-1 st.global kernel)' % shape, n, 'frames', ms, n * (13824 + shape[0] * shape[1] * 4))
        ms = timed(lambda: path.energy_heatmap(mfcc, True, shape[0], shape[1], want_energy=False, want_mask=False, out=out))
        report('energy + heat map %dx%d, one launch' % shape, n, 'frames', ms, n * (82944 + shape[0] * shape[1] * 4), 'aig_energy_heatmap')
        del out
        ms = timed(lambda: path.resize_mask(mask[:m], *shape))
        report('mask resize %dx%d' % shape, m, 'frames', ms, m * (1728 + shape[0] * shape[1]))
    heat = path.heatmap(energy[:m])
    frames_bgr = torch.randint(0, 256, (m, 224, 298, 3), device=dev, dtype=torch.uint8)
    ms = timed(lambda: path.overlay(heat, frames_bgr))
    report('jet overlay 224x298 (aig_overlay)', m, 'frames', ms, m * 224 * 298 * (4 + 3 + 3))
    thr = torch.linspace(0, 1, 101, device=dev, dtype=torch.float64)
    counts = torch.zeros(102, device=dev, dtype=torch.int64)
    half = n // 2
    ms = timed(lambda: path.iou_sweep(mask[:half], mask[half:2 * half], thr, pos=counts[:-1], num=counts[-1:]))
    report('IoU sweep, 101 thresholds', half, 'pairs', ms, half * 2 * 1728)
    boxes = [torch.from_numpy(v).to(dev) for v in synth.flickr_boxes(m, 3)]
    ms = timed(lambda: path.ciou_sweep(mask[:m], *boxes, thr, pos=counts[:-1], num=counts[-1:]))
    report('consensus IoU sweep 224x298, 101 thr', m, 'frames', ms, m * (1728 + 48), 'integer compute: 66 752 px/frame')
    from acoustic_image_generation_b200 import tables
    bank2 = aig.createfilters(256, 20, 300, 4000, 8000)
    dct2, lifter2, mfnorm2 = tables.mfcc_constants(20, 10, 22)
    p2 = aig.AcousticPath(0, stream=stream.cuda_stream, tables_=(bank2, dct2, lifter2, mfnorm2))
    rows2 = torch.randn((1 << 20, 256), device=dev, dtype=torch.float32).square_()
    ms = timed(lambda: p2.mfcc_rows(rows2))
    report('generic-table MFCC (256 bins, float64)', rows2.shape[0], 'rows', ms, rows2.shape[0] * (1024 + 40), 'fallback kernel, not tuned')
    p2.close()
    del rows2
    vec = torch.randn((n, 12), device=dev, dtype=torch.float32)
    ms = timed(lambda: path.tile_mfcc(vec, normalize=True))
    report('tile MFCC -> [36,48,12] (aig_tile_mfcc)', n, 'frames', ms, n * 82944, 'write-only')
    audio = torch.from_numpy(synth.audio_rows(64, 1, np.int32)).to(dev).repeat(64, 1)          # 4096 rows
    ms = timed(lambda: path.power_spectrum(audio))
    report('audio power spectrum (1024-pt rFFT f64)', audio.shape[0], 'rows', ms, audio.shape[0] * (4096 + 2048))
    ms = timed(lambda: path.build_spectrograms(audio))
    report('audio -> MFCC (_build_spectrograms)', audio.shape[0], 'rows', ms, audio.shape[0] * (4096 + 48))
    ms = timed(lambda: path.butter_lowpass_filter(audio))
    report('Butterworth filtfilt order 10', audio.shape[0], 'rows', ms, audio.shape[0] * 8192, 'sequential IIR, one thread per row')


if __name__ == '__main__':
    main()
