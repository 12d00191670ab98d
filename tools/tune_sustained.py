"""Ring-geometry choice under SUSTAINED load: the board sits at its power cap after ~1 s, the SM clock drops and the
consumer side of the pipeline gets slower, so the geometry that wins a 5-iteration burst need not win here.  Each
configuration runs back to back for --seconds and the last second's average is reported."""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import acoustic_image_generation_b200 as aig  # noqa: E402


def sustained(fn, stream, seconds, per_window=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.time()
    samples = []
    while time.time() - t0 < seconds:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(per_window):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        samples.append((time.time() - t0, e0.elapsed_time(e1) / per_window))
    first = samples[0][1]
    tail = [ms for t, ms in samples if t > seconds - 1.0]
    return first, sum(tail) / len(tail)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--frames', type=int, default=4096)
    ap.add_argument('--seconds', type=float, default=4.0)
    args = ap.parse_args()
    dev = torch.device('cuda', 0)
    stream = torch.cuda.current_stream()
    path = aig.AcousticPath(0, stream=stream.cuda_stream)
    n = args.frames
    power = torch.randn((n, 36, 48, 512), device=dev, dtype=torch.float32).square_()
    out = (torch.empty((n, 36, 48, 12), device=dev, dtype=torch.float32), torch.empty((n, 36, 48), device=dev, dtype=torch.float64),
           torch.empty((n, 36, 48), device=dev, dtype=torch.uint8))
    # warm the board up to its steady thermal / power state first
    sustained(lambda: path.mfcc_energy(power, flip=True, normalize_first=True, out=out), stream, 3.0)
    for v in (2, 1, 0, 2):
        path.set_option('fused_variant', v)
        first, steady = sustained(lambda: path.mfcc_energy(power, flip=True, normalize_first=True, out=out), stream, args.seconds)
        print('fused v%d: first window %.3f ms, steady %.3f ms  -> %.0f frames/s  %.0f GB/s' % (v, first, steady, n / steady * 1e3, n * 3628800 / steady / 1e6), flush=True)
    path.set_option('fused_variant', 2)
    rows_out = out[0].view(-1, 12)
    for v in (8, 5, 0, 3, 6, 8):
        path.set_option('mfcc_variant', v)
        first, steady = sustained(lambda: path.mfcc_rows(power, flip180=True, out=rows_out), stream, args.seconds)
        print('mfcc  v%d: first window %.3f ms, steady %.3f ms  -> %.0f frames/s  %.0f GB/s' % (v, first, steady, n / steady * 1e3, n * 3621888 / steady / 1e6), flush=True)


if __name__ == '__main__':
    main()
