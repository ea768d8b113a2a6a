"""Debug: per-phase cycle accounting of CTA 0 of the warp-specialised forward kernel (library built with `make TIMELINE=1`,
PWA_TIMELINE=1).  Usage: python tools/ws_profile.py [enc0|enc1|enc2|dec0] [s] [drop]"""
import os, sys
os.environ["PWA_TIMELINE"] = "1"
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pwa_b200
from pwa_b200 import functional as PF, _lib
dev = torch.device("cuda")
STAGES = {"enc0": (48, 4, (48, 48, 48)), "enc1": (96, 8, (24, 24, 24)), "enc2": (192, 16, (12, 12, 24)), "dec0": (192, 4, (12, 12, 24))}
stage = next((a for a in sys.argv[1:] if a in STAGES), "enc0")
C, heads, dims = STAGES[stage]
B, I, WS = 4, 64, (8, 8, 4)
shifted = "s" in sys.argv[1:]
drop = 0.1 if "drop" in sys.argv[1:] else 0.0
g = pwa_b200.get_geometry(dims, WS, (4, 4, 2) if shifted else (0, 0, 0))
qkv = torch.randn(B, g.P, g.N, 3 * C, device=dev).to(torch.bfloat16)
kvp = torch.randn(B, I, 2 * C, device=dev).to(torch.bfloat16)
th, tw, td = (0.3 * torch.randn(heads, w, w, device=dev) for w in WS)
tok = 0.3 * torch.randn(heads, I, device=dev)
ids = g.region_ids(dev) if g.masked else None
seed = torch.tensor([1, 2], dtype=torch.int32, device=dev) if drop else None
for _ in range(2):
    out = PF.prompted_window_attention_packed(qkv, kvp, th, tw, td, tok, ids, heads, WS, (C // heads) ** -0.5, PF.IMPL_TC, p_drop=drop, seed=seed)
torch.cuda.synchronize()
buf = np.zeros(64, dtype=np.int64)
_lib.lib.pwa_debug_fwd_timeline(buf.ctypes.data, buf.nbytes)
names = {0: ["wait operands", "window setup", "wait S", "exp pass", "st+arrive", "wait O", "epilogue"],
         8: ["wait operands", "window setup", "wait S", "exp pass", "st+arrive", "wait O", "epilogue"],
         16: ["wait free", "loads+stores", "barrier", "selectors+bound", "windows"],
         24: ["issue S (+operand wait)", "wait P", "wait O free", "issue PV"]}
for base, title in ((0, "softmax group 0 thread 0"), (8, "softmax group 1 thread 0"), (16, "staging thread 0"), (24, "issuer 0")):
    v = buf[base:base + 8]
    tot = sum(int(x) for i, x in enumerate(v[:len(names[base])]) if names[base][i] != "windows")
    print(title, "total", tot, {n: int(v[i]) for i, n in enumerate(names[base])})
