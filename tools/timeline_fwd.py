"""Debug: clock64 timeline of thread 0 / CTA 0 of the forward attention kernel (PWA_TIMELINE=1), one steady-state window."""
import os, sys
os.environ["PWA_TIMELINE"] = "1"
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pwa_b200
from pwa_b200 import functional as PF, _lib
dev = torch.device("cuda")
STAGES = {"enc0": (48, 4, (48, 48, 48)), "enc1": (96, 8, (24, 24, 24)), "enc2": (192, 16, (12, 12, 24))}
stage = next((a for a in sys.argv[1:] if a in STAGES), "enc0")
C, heads, dims = STAGES[stage]
B, I, WS = 4, 64, (8, 8, 4)
shifted = "s" in sys.argv[1:]
g = pwa_b200.get_geometry(dims, WS, (4, 4, 2) if shifted else (0, 0, 0))
qkv = torch.randn(B, g.P, g.N, 3 * C, device=dev).to(torch.bfloat16)
kvp = torch.randn(B, I, 2 * C, device=dev).to(torch.bfloat16)
th, tw, td = (0.3 * torch.randn(heads, w, w, device=dev) for w in WS)
tok = 0.3 * torch.randn(heads, I, device=dev)
ids = g.region_ids(dev) if g.masked else None
for _ in range(2):
    out = PF.prompted_window_attention_packed(qkv, kvp, th, tw, td, tok, ids, heads, WS, 12 ** -0.5, PF.IMPL_TC)
torch.cuda.synchronize()
buf = np.zeros(8002, dtype=np.int64)
n = _lib.lib.pwa_debug_fwd_timeline(buf.ctypes.data, buf.nbytes)
ev = [(int(buf[2 * i]), int(buf[2 * i + 1])) for i in range(2000) if buf[2 * i + 1] > 0]
starts = [i for i, e in enumerate(ev) if e[1] == 1]
print("events", len(ev), "windows", len(starts), "| setup clk", int(buf[8001] - buf[8000]), "| window starts / end after kernel entry",
      [e[0] - int(buf[8000]) for e in ev if e[1] in (1, 150)])
if len(starts) >= 3:
    seg = ev[starts[1]:starts[2] + 1]
    t0 = seg[0][0]
    print("window clk", seg[-1][0] - t0)
    print(" ".join(f"{tag}:{t - t0}" for t, tag in seg))
