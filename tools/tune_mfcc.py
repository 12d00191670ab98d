"""Time every ring geometry of the fused MFCC kernel on resident synthetic frames (run on a B200).

    python tools/tune_mfcc.py [--frames 4096] [--iters 5]

Prints one line per variant: ms per pass, frames/s, algorithmic GB/s (2048 B read + 48 B written per
spectrum) and the fraction of the measured HBM copy peak (MEASURED_PEAKS.json when present)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import acoustic_image_generation_b200 as aig  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--frames', type=int, default=4096)
    ap.add_argument('--iters', type=int, default=5)
    ap.add_argument('--variants', type=str, default='0,1,2,3,4,5,6,7,8,9')
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
    power = torch.randn((n, 36, 48, 512), device=dev, dtype=torch.float32).square_()
    out = torch.empty((n * 1728, 12), device=dev, dtype=torch.float32)
    bytes_per_pass = n * 1728 * (2048 + 48)
    for v, hint in [(int(x), hh) for hh in (0, 1) for x in args.variants.split(',')]:
        path.set_mfcc_variant(v)
        path.set_option('l2_evict_first', hint)
        for _ in range(2):
            path.mfcc_rows(power, out=out)
        torch.cuda.synchronize()
        times = []
        for _ in range(args.iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            path.mfcc_rows(power, out=out)
            e1.record(stream)
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = sorted(times)[len(times) // 2]
        gbs = bytes_per_pass / ms / 1e6
        print('variant %d hint %d: %.3f ms/pass (best %.3f)  %.0f frames/s  %.0f GB/s  %.3f of measured peak %.0f'
              % (v, hint, ms, min(times), n / ms * 1e3, gbs, gbs / peak, peak), flush=True)
    # energy stage on the MFCC images just produced
    img = out.view(n, 36, 48, 12)
    for _ in range(2):
        path.energy(img, normalize_first=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    path.energy(img, normalize_first=True)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print('energy stage: %.3f ms for %d frames (%.0f frames/s)' % (ms, n, n / ms * 1e3), flush=True)


if __name__ == '__main__':
    main()
