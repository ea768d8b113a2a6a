"""Debug: clock64 timeline of thread 0 / CTA 0 of the forward attention kernel (PWA_TIMELINE=1), one steady-state window."""
import os, sys
os.environ["PWA_TIMELINE"] = "1"
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pwa_b200
from pwa_b200 import functional as PF, _lib
dev = torch.device("cuda")
B, C, heads, I, WS = 4, 48, 4, 64, (8, 8, 4)
shifted = len(sys.argv) > 1 and sys.argv[1] == "s"
g = pwa_b200.get_geometry((48, 48, 48), WS, (4, 4, 2) if shifted else (0, 0, 0))
qkv = torch.randn(B, g.P, g.N, 3 * C, device=dev).to(torch.bfloat16)
kvp = torch.randn(B, I, 2 * C, device=dev).to(torch.bfloat16)
th, tw, td = (0.3 * torch.randn(heads, w, w, device=dev) for w in WS)
tok = 0.3 * torch.randn(heads, I, device=dev)
ids = g.region_ids(dev) if g.masked else None
for _ in range(2):
    out = PF.prompted_window_attention_packed(qkv, kvp, th, tw, td, tok, ids, heads, WS, 12 ** -0.5, PF.IMPL_TC)
torch.cuda.synchronize()
buf = np.zeros(4000, dtype=np.int64)
n = _lib.lib.pwa_debug_fwd_timeline(buf.ctypes.data, buf.nbytes)
ev = [(int(buf[2 * i]), int(buf[2 * i + 1])) for i in range(2000) if buf[2 * i + 1] > 0]
starts = [i for i, e in enumerate(ev) if e[1] == 1]
print("events", len(ev), "windows", len(starts))
if len(starts) >= 4:
    seg = ev[starts[2]:starts[3] + 1]
    t0 = seg[0][0]
    print("window clk", seg[-1][0] - t0)
    print(" ".join(f"{tag}:{t - t0}" for t, tag in seg))
