import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
import acoustic_image_generation_b200 as aig
from acoustic_image_generation_b200 import synth
p = aig.AcousticPath(0)
n = 128
arr = synth.power_frames(n, 0, 'chi2')
for t in (0, 1, 2, 4, 6, 8, 12, 16):
    p.set_option('host_copy_threads', t)
    for _ in range(2): p.mfcc_energy(arr, flip=True)
    t0 = time.perf_counter()
    for _ in range(6): p.mfcc_energy(arr, flip=True)
    dt = (time.perf_counter() - t0) / 6
    print('threads %2d: %7.2f ms  %6.1f GB/s' % (t, dt * 1e3, n * 3538944 / dt / 1e9))
# plain memcpy scaling on this host
import threading
src = np.frombuffer(arr, dtype=np.uint8)
dst = np.empty_like(src)
for t in (1, 2, 4, 8, 16):
    parts = np.array_split(np.arange(len(src)), 1)  # placeholder
    step = len(src) // t
    def work(k):
        np.copyto(dst[k * step:(k + 1) * step], src[k * step:(k + 1) * step])
    for rep in range(2):
        th = [threading.Thread(target=work, args=(k,)) for k in range(t)]
        t0 = time.perf_counter()
        [x.start() for x in th]; [x.join() for x in th]
        dt = time.perf_counter() - t0
    print('numpy memcpy %2d threads: %6.1f GB/s' % (t, len(src) / dt / 1e9))
