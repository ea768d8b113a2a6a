"""Launch partition / reverse a few times at the BASELINE stage shapes (for `ncu --metrics gpu__time_duration.sum`)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pwa_b200
from pwa_b200 import functional as PF
shapes = [(48, (48, 48, 48)), (96, (24, 24, 24)), (192, (12, 12, 24))]
if len(sys.argv) > 1:
    shapes = shapes[:int(sys.argv[1])]
for C, dims in shapes:
    for shift in ((0, 0, 0), (4, 4, 2)):
        g = pwa_b200.get_geometry(dims, (8, 8, 4), shift)
        for dt in (torch.bfloat16, torch.float32):
            xs = [torch.randn(4, C, *dims, device="cuda").to(dt) for _ in range(3)]
            for i in range(3):
                t = PF._partition_raw(xs[i], g, 0)
                y = PF._reverse_raw(t, g, 1)
                z = PF.reverse_add_tokens(t, t, g)
torch.cuda.synchronize()
print("ok")
