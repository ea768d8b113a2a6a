"""Debug: merged clock64 timeline of CTA 0 of the backward attention kernel (PWA_TIMELINE=1): compute group 0 / 1
(thread 0 of each), S / V / Q issuers and the staging warp, for one steady-state window."""
import os, sys
os.environ["PWA_TIMELINE"] = "1"
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pwa_b200
from pwa_b200 import functional as PF
dev = torch.device("cuda")
STAGES = {"enc0": (48, 4, (48, 48, 48)), "enc1": (96, 8, (24, 24, 24)), "enc2": (192, 16, (12, 12, 24))}
stage = next((a for a in sys.argv[1:] if a in STAGES), "enc0")
C, heads, dims = STAGES[stage]
B, I, WS = 4, 64, (8, 8, 4)
shifted = "s" in sys.argv[1:]
g = pwa_b200.get_geometry(dims, WS, (4, 4, 2) if shifted else (0, 0, 0))
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
names = {0: "cA", 8192: "cB", 2048: "S", 4096: "V", 10240: "Q", 6144: "P"}
allev = []
for base, nm in names.items():
    ev = [(int(d[base + 2 * i]), int(d[base + 2 * i + 1])) for i in range(1000)]
    allev += [(t, nm, tag) for t, tag in ev if 0 < tag < 200 and t > 0]
ks, ke = int(d[16000]), int(d[16001])
firsts = {nm: min(t for t, n2, tag in allev if n2 == nm) for nm in names.values() if any(n2 == nm for _, n2, _ in allev)}
lasts = {nm: [(t - ks, tag) for t, n2, tag in sorted(allev) if n2 == nm][-8:] for nm in names.values()}
print("setup clk", ke - ks, "| first event per actor (clk after kernel entry)", {k: v - ks for k, v in firsts.items()})
print("last events per actor", lasts)
ca = [(t, tag) for t, nm, tag in allev if nm == "cA"]
starts = [t for t, tag in ca if tag == 1]
if len(starts) < 4:
    print("not enough events", len(ca)); sys.exit(0)
t0, t1 = starts[2], starts[3]
print("window clk", t1 - t0)
for t, nm, tag in sorted(allev):
    if t0 - 200 <= t <= t1 + 200:
        print(f"{t - t0:7d} {nm:>3} {tag}")
