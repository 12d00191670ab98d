"""Explore the chained MFCC -> energy pipeline on a B200: ring variant x chunk size x overlap.

    python tools/tune_chain.py [--frames 4096] [--iters 5]
"""
import argparse
import itertools
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
    ap.add_argument('--variants', default='5,8')
    ap.add_argument('--chunks', default='256,512,1024')
    ap.add_argument('--energy-ctas', default='1,2,3')
    args = ap.parse_args()
    dev = torch.device('cuda', 0)
    stream = torch.cuda.current_stream()
    path = aig.AcousticPath(0, stream=stream.cuda_stream)
    n = args.frames
    power = torch.randn((n, 36, 48, 512), device=dev, dtype=torch.float32).square_()
    out = (torch.empty((n, 36, 48, 12), device=dev, dtype=torch.float32),
           torch.empty((n, 36, 48), device=dev, dtype=torch.float64),
           torch.empty((n, 36, 48), device=dev, dtype=torch.uint8))
    for variant, overlap, chunk, ectas in itertools.product([int(v) for v in args.variants.split(',')], (1,),
                                                            [int(c) for c in args.chunks.split(',')],
                                                            [int(c) for c in args.energy_ctas.split(',')]):
        path.set_option('chain_energy_ctas_per_sm', ectas)
        path.set_option('mfcc_variant', variant)
        path.set_option('chain_overlap', overlap)
        path.set_option('chain_chunk_frames', chunk)
        for _ in range(2):
            path.mfcc_energy(power, flip=True, normalize_first=True, out=out)
        torch.cuda.synchronize()
        path.set_option('profile', 1)
        times = []
        for _ in range(args.iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            path.mfcc_energy(power, flip=True, normalize_first=True, out=out)
            e1.record(stream)
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        prof = path.profile_read()
        path.set_option('profile', 0)
        ms = sorted(times)[len(times) // 2]
        print('variant %d overlap %d chunk %5d ectas %d: %.3f ms/pass  %.0f frames/s  | mfcc kernels %.3f ms, energy kernels %.3f ms per pass'
              % (variant, overlap, chunk, ectas, ms, n / ms * 1e3, prof['mfcc'][0] / args.iters, prof['energy'][0] / args.iters), flush=True)


if __name__ == '__main__':
    main()
