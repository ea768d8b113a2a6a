import subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, torch
sys.path.insert(0, %r)
import pwa_b200
from pwa_b200 import functional as PF
dims, shift, C, dt = eval(sys.argv[1]), eval(sys.argv[2]), int(sys.argv[3]), getattr(torch, sys.argv[4])
g = pwa_b200.get_geometry(dims, (8, 8, 4), shift)
x = torch.randn(2, C, *dims, device="cuda").to(dt)
a = PF._partition_raw(x, g, 0); torch.cuda.synchronize()
print("ran; pads", g.pads, "sp", g.sp, flush=True)
''' % ROOT
for dbg in ("0", "1", "2", "3"):
    for c in [((16, 12, 16), (4, 4, 2), 12, "float32"), ((16, 12, 16), (0, 0, 0), 12, "float32")]:
        r = subprocess.run([sys.executable, "-c", CHILD, *map(str, c)], capture_output=True, text=True, timeout=120,
                           env=dict(os.environ, CUDA_LAUNCH_BLOCKING="1", PWA_TMA_DEBUG=dbg))
        err = [l for l in r.stderr.strip().splitlines() if "rror" in l][-1:] if r.returncode else []
        print("debug", dbg, c, "->", r.stdout.strip(), err)
