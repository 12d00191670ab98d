import os, sys
sys.path.insert(0, os.getcwd())
import torch
import acoustic_image_generation_b200 as aig
p = aig.AcousticPath(0)
for n in (1, 4, 16, 64, 148, 296):
    power = torch.rand(n, 36, 48, 512, device='cuda') ** 2
    out = (torch.empty(n, 36, 48, 12, device='cuda'), torch.empty(n, 36, 48, device='cuda', dtype=torch.float64), torch.empty(n, 36, 48, device='cuda', dtype=torch.uint8))
    res = []
    for mode in (2, 0):
        p.set_option('chain_mode', mode)
        for _ in range(5): p.mfcc_energy(power, flip=True, out=out)
        torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50): p.mfcc_energy(power, flip=True, out=out)
        b.record(); torch.cuda.synchronize()
        res.append(a.elapsed_time(b) / 50 * 1e3)
    p.set_option('chain_mode', 2)
    print('%4d frames: fused %8.1f us   two kernels %8.1f us' % (n, res[0], res[1]))
