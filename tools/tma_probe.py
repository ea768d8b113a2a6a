"""Debug: run the TMA partition / reverse kernels on a matrix of geometries, each in its own process (CUDA errors are
sticky), and compare with the generic kernel."""
import subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, torch
sys.path.insert(0, %r)
import pwa_b200
from pwa_b200 import functional as PF
dims, shift, C, dt = eval(sys.argv[1]), eval(sys.argv[2]), int(sys.argv[3]), getattr(torch, sys.argv[4])
g = pwa_b200.get_geometry(dims, (8, 8, 4), shift)
torch.manual_seed(0)
x = torch.randn(2, C, *dims, device="cuda").to(dt)
for lo in (0, 1):
    a = PF._partition_raw(x, g, lo); torch.cuda.synchronize()
    b = PF._partition_raw(x, g, lo, force_generic=True)
    print("part lo", lo, "equal", torch.equal(a, b), flush=True)
    t = torch.randn(2, g.P, g.N, C, device="cuda").to(dt)
    a = PF._reverse_raw(t, g, lo); torch.cuda.synchronize()
    b = PF._reverse_raw(t, g, lo, force_generic=True)
    print("rev  lo", lo, "equal", torch.equal(a, b), flush=True)
''' % ROOT
cases = [((16, 16, 16), (0, 0, 0), 12, "float32"), ((16, 16, 16), (4, 4, 2), 48, "bfloat16"), ((12, 16, 16), (4, 4, 2), 12, "float32"),
         ((16, 12, 16), (4, 4, 2), 12, "float32"), ((16, 16, 12), (4, 4, 2), 12, "float32"), ((12, 12, 8), (4, 4, 2), 12, "float32"),
         ((12, 12, 24), (4, 4, 2), 48, "bfloat16"), ((12, 12, 24), (4, 4, 2), 192, "bfloat16"), ((24, 24, 24), (4, 4, 2), 96, "bfloat16")]
for c in cases:
    r = subprocess.run([sys.executable, "-c", CHILD, *map(str, c)], capture_output=True, text=True, timeout=120,
                       env=dict(os.environ, CUDA_LAUNCH_BLOCKING="1"))
    out = " | ".join(r.stdout.strip().splitlines())
    err = [l for l in r.stderr.strip().splitlines() if "Error" in l][-1:] if r.returncode else []
    print(c, "->", out, err)
