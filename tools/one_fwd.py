"""One shape of the attention forward, a few launches (for ncu).  Usage: python tools/one_fwd.py [enc0] [s] [drop] [bwd]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pwa_b200
from pwa_b200 import functional as PF
dev = torch.device("cuda")
STAGES = {"enc0": (48, 4, (48, 48, 48)), "enc1": (96, 8, (24, 24, 24)), "enc2": (192, 16, (12, 12, 24)), "dec0": (192, 4, (12, 12, 24))}
stage = next((a for a in sys.argv[1:] if a in STAGES), "enc0")
C, heads, dims = STAGES[stage]
B, I, WS = 4, 64, (8, 8, 4)
shifted, drop, bwd = "s" in sys.argv[1:], (0.1 if "drop" in sys.argv[1:] else 0.0), "bwd" in sys.argv[1:]
g = pwa_b200.get_geometry(dims, WS, (4, 4, 2) if shifted else (0, 0, 0))
torch.manual_seed(0)
qkv = torch.randn(B, g.P, g.N, 3 * C, device=dev).to(torch.bfloat16).requires_grad_(bwd)
kvp = torch.randn(B, I, 2 * C, device=dev).to(torch.bfloat16).requires_grad_(bwd)
th, tw, td = (0.3 * torch.randn(heads, w, w, device=dev) for w in WS)
tok = 0.3 * torch.randn(heads, I, device=dev)
ids = g.region_ids(dev) if g.masked else None
seed = torch.tensor([1, 2], dtype=torch.int32, device=dev) if drop else None
for _ in range(3):
    out = PF.prompted_window_attention_packed(qkv, kvp, th, tw, td, tok, ids, heads, WS, (C // heads) ** -0.5, PF.IMPL_TC, p_drop=drop, seed=seed)
    if bwd:
        out.backward(torch.ones_like(out))
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
