"""Burst vs sustained: run the fused MFCC+energy pass (and, for context, a plain device-to-device copy of the same
bytes) back to back for several seconds and print the throughput of each half-second window with the SM clock / power
NVML reports - does the HBM stream slow down once the board sits at its power cap?"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import acoustic_image_generation_b200 as aig  # noqa: E402

try:
    import pynvml
    pynvml.nvmlInit()
    NV = pynvml.nvmlDeviceGetHandleByIndex(0)
except Exception:
    NV = None


def nv():
    if NV is None:
        return ''
    return 'sm %d MHz mem %d MHz %.0f W %d C' % (pynvml.nvmlDeviceGetClockInfo(NV, pynvml.NVML_CLOCK_SM),
                                                 pynvml.nvmlDeviceGetClockInfo(NV, pynvml.NVML_CLOCK_MEM),
                                                 pynvml.nvmlDeviceGetPowerUsage(NV) / 1e3,
                                                 pynvml.nvmlDeviceGetTemperature(NV, pynvml.NVML_TEMPERATURE_GPU))


def windows(name, fn, unit_bytes, seconds=6.0, per_window=25):
    stream = torch.cuda.current_stream()
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t_end = time.time() + seconds
    while time.time() < t_end:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(per_window):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print('%-28s %.3f ms/pass  %.0f GB/s   %s' % (name, ms / per_window, unit_bytes * per_window / ms / 1e6, nv()), flush=True)


def main():
    dev = torch.device('cuda', 0)
    path = aig.AcousticPath(0, stream=torch.cuda.current_stream().cuda_stream)
    n = 4096
    power = torch.randn((n, 36, 48, 512), device=dev, dtype=torch.float32).square_()
    out = (torch.empty((n, 36, 48, 12), device=dev, dtype=torch.float32), torch.empty((n, 36, 48), device=dev, dtype=torch.float64),
           torch.empty((n, 36, 48), device=dev, dtype=torch.uint8))
    print('idle:', nv())
    windows('fused MFCC+energy', lambda: path.mfcc_energy(power, flip=True, normalize_first=True, out=out), n * 3628800)
    time.sleep(2.0)
    dst = torch.empty_like(power[:n // 2])
    src = power[:n // 2]
    windows('copy (read+write bytes)', lambda: dst.copy_(src), 2 * src.numel() * 4)
    time.sleep(2.0)
    windows('read-only sum (torch)', lambda: torch.sum(power), power.numel() * 4, seconds=4.0, per_window=10)


if __name__ == '__main__':
    main()
