"""Heat-map kernel throughput against the frame count (resident float64 energies in, float32 images out)."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
import acoustic_image_generation_b200 as aig
p = aig.AcousticPath(0)
for n in (2048, 8192, 16384):
    e = torch.rand(n, 36, 48, device='cuda', dtype=torch.float64)
    for hw in ((224, 298), (224, 224)):
        for _ in range(3): out = p.heatmap(e, *hw)    # same allocation pattern as the timed loop
        torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): out = p.heatmap(e, *hw)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        print(n, hw, '%.3f ms  %.1f M frames/s  %.0f GB/s' % (ms, n / ms / 1e3, n * (hw[0] * hw[1] * 4 + 1728 * 8) / ms / 1e6))
        del out
