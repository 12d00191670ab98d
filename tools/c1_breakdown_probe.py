"""Where the time of a 16-frame host call (BASELINE configs[0]) goes: device-resident, pinned and pageable inputs, next
to a bare pinned cudaMemcpy of the same bytes.  Run on a GPU box: python tools/c1_breakdown_probe.py"""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
import torch
import acoustic_image_generation_b200 as aig
from acoustic_image_generation_b200 import synth

p = aig.AcousticPath(0)
n = 16
power = synth.power_frames(n, 0, 'chi2')


def lat(fn, reps=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    ts.sort()
    return ts[len(ts) // 2] * 1e3, ts[0] * 1e3


d_power = torch.from_numpy(power).cuda()
d_out = (torch.empty(n, 36, 48, 12, device='cuda'), torch.empty(n, 36, 48, device='cuda', dtype=torch.float64),
         torch.empty(n, 36, 48, device='cuda', dtype=torch.uint8))
h_pin = torch.empty(power.shape, dtype=torch.float32, pin_memory=True); h_pin.copy_(torch.from_numpy(power))
pin_out = tuple(torch.empty(t.shape, dtype=t.dtype, pin_memory=True).numpy() for t in d_out)
pg_out = tuple(np.empty(t.shape, dtype=o.dtype) for t, o in zip(d_out, pin_out))
np_pin = h_pin.numpy()
print('bytes in %.1f MB' % (power.nbytes / 1e6))
print('device in / device out            median %.3f ms  min %.3f ms' % lat(lambda: p.mfcc_energy(d_power, flip=True, normalize_first=True, out=d_out)))
print('bare pinned H2D copy              median %.3f ms  min %.3f ms' % lat(lambda: d_power.copy_(h_pin, non_blocking=True)))
print('pinned in / pinned out            median %.3f ms  min %.3f ms' % lat(lambda: p.mfcc_energy(np_pin, flip=True, normalize_first=True, out=pin_out)))
print('pageable in / pinned out          median %.3f ms  min %.3f ms' % lat(lambda: p.mfcc_energy(power, flip=True, normalize_first=True, out=pin_out)))
print('pageable in / pageable out (prealloc) median %.3f ms  min %.3f ms' % lat(lambda: p.mfcc_energy(power, flip=True, normalize_first=True, out=pg_out)))
print('pageable in / fresh arrays out    median %.3f ms  min %.3f ms' % lat(lambda: p.mfcc_energy(power, flip=True, normalize_first=True)))
scratch = np.empty_like(power)
print('host memcpy of the input, 1 thread  median %.3f ms  min %.3f ms' % lat(lambda: np.copyto(scratch, power)))
big = synth.power_frames(64, 1, 'chi2')
big = np.concatenate([big] * 4, 0)                      # 256 frames, 906 MB pageable
big_out = tuple(np.empty((256,) + o.shape[1:], dtype=o.dtype) for o in pg_out)
for streaming in (0, 1):
    p.set_option('host_copy_streaming', streaming)
    for threads in (1, 2, 3, 4, 6, 8):
        p.set_option('host_copy_threads', threads)
        a = lat(lambda: p.mfcc_energy(power, flip=True, normalize_first=True, out=pg_out))
        b = lat(lambda: p.mfcc_energy(big, flip=True, normalize_first=True, out=big_out), reps=5)
        print('streaming fill %d, host_copy_threads %2d: 16 pageable frames median %.3f ms min %.3f ms | 256 frames median %.2f ms = %.1f k frames/s'
              % (streaming, threads, a[0], a[1], b[0], 256 / b[0]))
# the same 16-frame call with a COLD source: sixteen different arrays in rotation (906 MB, more than any host cache) -
# what a pipeline that produces a fresh batch per call sees; the loop above re-sends one array, which stays cached
cold = [big[16 * i:16 * i + 16].copy() for i in range(16)]
state = {'i': 0}
def cold_call():
    state['i'] = (state['i'] + 1) % 16
    p.mfcc_energy(cold[state['i']], flip=True, normalize_first=True, out=pg_out)
for streaming in (0, 1):
    p.set_option('host_copy_streaming', streaming)
    for threads in (3, 4, 6):
        p.set_option('host_copy_threads', threads)
        print('cold source, streaming fill %d, host_copy_threads %d: 16 pageable frames median %.3f ms min %.3f ms'
              % ((streaming, threads) + lat(cold_call, reps=48)))
p.set_option('host_copy_threads', -1)
print('cpu_count', os.cpu_count(), 'affinity', len(os.sched_getaffinity(0)))
