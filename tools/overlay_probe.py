import os, sys
sys.path.insert(0, os.getcwd())
import torch
import acoustic_image_generation_b200 as aig
p = aig.AcousticPath(0)
def timed(fn, reps=9):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts=[]
    for _ in range(reps):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts)//2]
for n in (2048, 4096, 8192):
    heat = torch.rand(n, 224, 298, device='cuda')
    bgr = torch.randint(0, 256, (n, 224, 298, 3), device='cuda', dtype=torch.uint8)
    out = torch.empty(n, 224, 298, 3, device='cuda', dtype=torch.uint8)
    lut = aig.tables.jet_lut()
    ms = timed(lambda: p._check(p._lib.aig_overlay(p._h, heat.data_ptr(), bgr.data_ptr(), n, 224, 298, 0.7, lut.ctypes.data, out.data_ptr())))
    print('overlay %5d frames  %.3f ms  %.2f M frames/s  %.2f TB/s (heat + BGR in, RGB out)' % (n, ms, n/ms/1e3, n*224*298*10/ms/1e9))
    p.set_option('overlay_luma', 0)
    out2 = torch.empty_like(out)
    ms0 = timed(lambda: p._check(p._lib.aig_overlay(p._h, heat.data_ptr(), bgr.data_ptr(), n, 224, 298, 0.7, lut.ctypes.data, out2.data_ptr())))
    p.set_option('overlay_luma', 1)
    print('   BGR read in both passes (overlay_luma = 0)  %.3f ms  %.2f M frames/s; outputs equal: %s' % (ms0, n/ms0/1e3, bool(torch.equal(out, out2))))
    del out2
    smooth = torch.nn.functional.interpolate(torch.rand(n, 1, 9, 12, device='cuda'), size=(224, 298), mode='bilinear').squeeze(1).contiguous()
    ms = timed(lambda: p._check(p._lib.aig_overlay(p._h, smooth.data_ptr(), bgr.data_ptr(), n, 224, 298, 0.7, lut.ctypes.data, out.data_ptr())))
    print('   smooth heat maps    %.3f ms  %.2f M frames/s' % (ms, n/ms/1e3))
    del heat, bgr, out, smooth
