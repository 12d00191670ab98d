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
    ap.add_argument('--variants', default='0')
    ap.add_argument('--chunks', default='512')
    ap.add_argument('--energy-ctas', default='3')
    args = ap.parse_args()
    dev = torch.device('cuda', 0)
    stream = torch.cuda.current_stream()
    path = aig.AcousticPath(0, stream=stream.cuda_stream)
    n = args.frames
    power = torch.randn((n, 36, 48, 512), device=dev, dtype=torch.float32).square_()
    out = (torch.empty((n, 36, 48, 12), device=dev, dtype=torch.float32),
           torch.empty((n, 36, 48), device=dev, dtype=torch.float64),
           torch.empty((n, 36, 48), device=dev, dtype=torch.uint8))
    configs = [('fused v%d' % v, {'chain_mode': 2, 'fused_variant': v}) for v in (2, 1, 0)]
    configs += [('overlap v%d chunk %d ectas %d' % (v, c, e),
                 {'chain_mode': 1, 'mfcc_variant': v, 'chain_chunk_frames': c, 'chain_energy_ctas_per_sm': e})
                for v in [int(x) for x in args.variants.split(',')] for c in [int(x) for x in args.chunks.split(',')]
                for e in [int(x) for x in args.energy_ctas.split(',')]]
    configs += [('sequential v5 chunk 4096', {'chain_mode': 0, 'mfcc_variant': 5, 'chain_chunk_frames': 4096})]
    for name, opts in configs:
        for k, v in opts.items():
            path.set_option(k, v)
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
        print('%-36s %.3f ms/pass (best %.3f)  %.0f frames/s  %.0f GB/s | mfcc/fused kernels %.3f ms, energy kernels %.3f ms per pass'
              % (name, ms, min(times), n / ms * 1e3, n * 3628800 / ms / 1e6, prof['mfcc'][0] / args.iters,
                 prof['energy'][0] / args.iters), flush=True)


if __name__ == '__main__':
    main()
