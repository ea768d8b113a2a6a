"""Debug: per-unit clock64 timeline of CTA 0 of the backward attention kernel (PWA_TIMELINE=1)."""
import os, sys
os.environ["PWA_TIMELINE"] = "1"
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pwa_b200
from pwa_b200 import functional as PF
dev = torch.device("cuda")
B, C, heads, I, WS = 4, 48, 4, 64, (8, 8, 4)
shifted = len(sys.argv) > 1
g = pwa_b200.get_geometry((48, 48, 48), WS, (4, 4, 2) if shifted else (0, 0, 0))
qkv = torch.randn(B, g.P, g.N, 3 * C, device=dev).to(torch.bfloat16).requires_grad_(True)
kvp = torch.randn(B, I, 2 * C, device=dev).to(torch.bfloat16).requires_grad_(True)
th, tw, td = (0.3 * torch.randn(heads, w, w, device=dev) for w in WS)
tok = 0.3 * torch.randn(heads, I, device=dev)
ids = g.region_ids(dev) if g.masked else None
for _ in range(2):
    out = PF.prompted_window_attention_packed(qkv, kvp, th, tw, td, tok, ids, heads, WS, 12 ** -0.5, PF.IMPL_TC)
    out.backward(torch.randn_like(out))
torch.cuda.synchronize()
d = PF._WindowAttentionPacked.last_delta.reshape(-1).view(torch.int64).cpu()
for name, base in (("compute tid0", 0), ("S issuer", 2048), ("V issuer", 4096), ("producer", 6144)):
    ev = [(int(d[base + 2 * i]), int(d[base + 2 * i + 1])) for i in range(1000)]
    ev = [e for e in ev if 0 < e[1] < 200]
    # second window of this CTA: from the 2nd tag==1 to the 3rd
    starts = [i for i, e in enumerate(ev) if e[1] == 1]
    if len(starts) < 3:
        print(name, "not enough events", len(ev)); continue
    seg = ev[starts[1]:starts[2] + 1]
    t0 = seg[0][0]
    print(name, "window total clk", seg[-1][0] - t0)
    print(" ".join(f"{tag}:{t - t0}" for t, tag in seg))
