"""One multi-window case of the tcgen05 attention kernels for compute-sanitizer (enc1 shape, B = 1: 432 window-heads on
144 CTAs = 3 windows per CTA; shifted, prompts, dropout) plus the token-domain kernels of one block.
    compute-sanitizer --tool memcheck|racecheck python tools/sanitize_case.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pwa_b200
from pwa_b200 import functional as PF
dev = torch.device("cuda")
torch.manual_seed(0)
C, heads, dims, WS, I = 96, 8, (24, 24, 24), (8, 8, 4), 64
g = pwa_b200.get_geometry(dims, WS, (4, 4, 2))
qkv = torch.randn(1, g.P, g.N, 3 * C, device=dev).bfloat16().requires_grad_(True)
kvp = torch.randn(1, I, 2 * C, device=dev).bfloat16().requires_grad_(True)
th, tw, td = (0.3 * torch.randn(heads, w, w, device=dev) for w in WS)
tok = 0.3 * torch.randn(heads, I, device=dev)
ids = g.region_ids(dev)
seed = torch.tensor([11, 22], dtype=torch.int32, device=dev)
for p_drop in (0.0, 0.1):
    out = PF.prompted_window_attention_packed(qkv, kvp, th, tw, td, tok, ids, heads, WS, (C // heads) ** -0.5, PF.IMPL_TC,
                                              p_drop=p_drop, seed=seed if p_drop else None)
    out.backward(torch.ones_like(out))
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all() and torch.isfinite(qkv.grad.float()).all()
blk = pwa_b200.SwinTransformerBlock(hidden_channels=48, window_size=WS, pos_bias_embed_dim=64, num_heads=4, max_prompts=1,
                                    tokens_per_prompt=64, shift_size=(4, 4, 2), attn_drop=0.1, proj_drop=0.1).to(dev).train()
x = torch.randn(1, 48, 16, 16, 8, device=dev).bfloat16().requires_grad_(True)
y = blk(x, 0.2 * torch.randn(1, 64, 48, device=dev).bfloat16())
y.float().square().mean().backward()
torch.cuda.synchronize()
print("sanitize case OK", float(out.float().abs().mean()), float(y.float().abs().mean()))
