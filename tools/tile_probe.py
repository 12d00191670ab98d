"""aig_tile_mfcc by frame count:  python tools/tile_probe.py"""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
import acoustic_image_generation_b200 as aig
p = aig.AcousticPath(0)
for n in (4096, 16384, 65536):
    v = torch.randn(n + 1, 12, device='cuda')
    out = torch.empty(n * 20736 + 4, device='cuda')
    for name, vin, o in (('st.global.v4', v, out),):
        ts = []
        for i in range(12):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            p._check(p._lib.aig_tile_mfcc(p._h, vin.data_ptr(), n, 0, o.data_ptr()))
            b.record(); torch.cuda.synchronize()
            if i >= 3: ts.append(a.elapsed_time(b))
        ms = sorted(ts)[len(ts)//2]
        print('tile %-10s %6d frames %.3f ms %.1f M frames/s %.2f TB/s' % (name, n, ms, n/ms/1e3, n*82944/ms/1e9), flush=True)
    del v, out
