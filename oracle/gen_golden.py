"""Generate tests/golden/*.npz by EXECUTING THE LIVE REFERENCE (test infrastructure only).

Run in the build container, where /root/reference is mounted:
    python -m oracle.gen_golden
The reference has no tests or golden vectors of its own, so these files
are what pins the oracle restatement and the CUDA path.  Each block case stores the
fp32 inputs + state dict and the reference's output and all gradients computed in
float64 (the reference block runs unmodified in .double()).  Loss for the gradients is
sum(out * g) with a stored upstream gradient g.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name, C, heads, ws, dims, shift (None = unshifted), I (0 = no prompts), E, batch
BLOCK_CASES = [
    ("blk_small_noshift",     12, 4, (4, 4, 2), (8, 8, 4),   None,      8, 16, 2),
    ("blk_small_shift",       12, 4, (4, 4, 2), (8, 8, 4),   (2, 2, 1), 8, 16, 2),
    ("blk_small_shift_nop",   12, 4, (4, 4, 2), (8, 8, 4),   (2, 2, 1), 0, 16, 2),
    ("blk_pad_shift",         12, 2, (4, 4, 2), (6, 6, 6),   (2, 2, 1), 8, 16, 1),
    ("blk_pad_divisible_ax",  12, 2, (4, 4, 2), (6, 8, 4),   (2, 2, 1), 8, 16, 1),
    ("blk_pad_odd",           12, 2, (4, 4, 2), (5, 9, 3),   (2, 2, 1), 8, 16, 1),
    ("blk_axis_le_window",    12, 4, (4, 4, 2), (8, 8, 2),   (2, 2, 1), 8, 16, 1),
    ("blk_w884_dh12_shift",   48, 4, (8, 8, 4), (16, 16, 8), (4, 4, 2), 64, 64, 1),
    ("blk_w884_dh12_noshift", 48, 4, (8, 8, 4), (8, 8, 8),   None,      64, 64, 1),
    ("blk_w884_dh24",         48, 2, (8, 8, 4), (16, 8, 8),  (4, 4, 2), 64, 64, 1),
    ("blk_w884_dh48_pad",     96, 2, (8, 8, 4), (6, 6, 12),  (4, 4, 2), 64, 64, 1),
    ("blk_w884_dh3",          12, 4, (8, 8, 4), (16, 16, 8), (4, 4, 2), 64, 64, 1),
    ("blk_w884_dh6_h16",      96, 16, (8, 8, 4), (8, 8, 8),  (4, 4, 2), 64, 64, 1),
]

# name, dims, ws, shift, pads-from-dims?   (mask + index-map cases, bit-exact)
GEOM_CASES = [
    ("geo_div_shift",      (8, 8, 4),    (4, 4, 2), (2, 2, 1)),
    ("geo_pad_all",        (6, 6, 6),    (4, 4, 2), (2, 2, 1)),
    ("geo_pad_divisible",  (6, 8, 4),    (4, 4, 2), (2, 2, 1)),
    ("geo_axis_le_window", (8, 8, 2),    (4, 4, 2), (2, 2, 1)),
    ("geo_w884_24",        (24, 24, 24), (8, 8, 4), (4, 4, 2)),
    ("geo_w884_12x12x24",  (12, 12, 24), (8, 8, 4), (4, 4, 2)),
    ("geo_w884_noshift",   (16, 16, 8),  (8, 8, 4), (0, 0, 0)),
    ("geo_odd",            (5, 9, 3),    (4, 4, 2), (2, 2, 1)),
]


def _rand_like_ref_inputs(gen, shape, scale=1.0):
    return (torch.randn(shape, generator=gen) * scale).float()


def make_block(ref, C, heads, ws, shift, I, E, seed):
    torch.manual_seed(seed)
    blk = ref.SwinTransformerBlock(
        hidden_channels=C, window_size=ws, pos_bias_embed_dim=E, num_heads=heads,
        max_prompts=1, tokens_per_prompt=max(I, 1), use_token_params=I > 0,
        shift_size=shift if shift is not None else (0, 0, 0))
    # make LN / bias parameters non-trivial so their gradients are exercised
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, prm in blk.named_parameters():
            if n.endswith("norm.weight"):
                prm.add_(0.2 * torch.randn(prm.shape, generator=g))
            elif n.endswith("norm.bias") or n.endswith(".bias"):
                prm.add_(0.1 * torch.randn(prm.shape, generator=g))
            elif n.startswith("pe."):
                prm.mul_(3.0)          # make position bias matter
    return blk


def run_block_case(ref, case, seed):
    name, C, heads, ws, dims, shift, I, E, B = case
    blk = make_block(ref, C, heads, ws, shift, I, E, seed)
    g = torch.Generator().manual_seed(seed + 7)
    x = _rand_like_ref_inputs(g, (B, C, *dims))
    p = _rand_like_ref_inputs(g, (B, I, C), 0.5) if I > 0 else None
    go = _rand_like_ref_inputs(g, (B, C, *dims))
    out = {"x": x.numpy(), "go": go.numpy()}
    if p is not None:
        out["p"] = p.numpy()
    sd = {k: v.detach().clone() for k, v in blk.state_dict().items()}
    for k, v in sd.items():
        out["sd." + k] = v.numpy()

    blk64 = blk.double()
    x64 = x.double().requires_grad_(True)
    p64 = p.double().requires_grad_(True) if p is not None else None
    y = blk64(x64, p64)
    (y * go.double()).sum().backward()
    # computed in float64, stored rounded to float32 (6e-8 relative: far below every tolerance)
    f32 = lambda t: t.detach().to(torch.float32).numpy()
    out["out"] = f32(y)
    out["grad.x"] = f32(x64.grad)
    if p is not None:
        out["grad.p"] = f32(p64.grad)
    for n, prm in blk64.named_parameters():
        out["grad." + n] = f32(prm.grad) if prm.grad is not None else np.zeros(prm.shape, np.float32)
    out["meta"] = np.array([C, heads, *ws, *dims, *(shift or (0, 0, 0)), I, E, B], dtype=np.int64)
    return out


def run_geom_case(ref, case):
    name, dims, ws, shift_cfg = case
    blk = ref.SwinTransformerBlock(hidden_channels=4, window_size=ws, pos_bias_embed_dim=4,
                                   num_heads=1, max_prompts=1, tokens_per_prompt=1, shift_size=shift_cfg)
    shift = blk.get_shift_size(dims)
    import math
    pads = (0, 0, 0, 0, 0, 0)
    if any(d % w for d, w in zip(dims, ws)):
        pads = []
        for d, w in zip(dims, ws):
            pads += [math.floor((w - d % w) / 2), math.ceil((w - d % w) / 2)]
        pads = tuple(pads)
    sp = tuple(d + pads[2 * a] + pads[2 * a + 1] for a, d in enumerate(dims))
    out = {"meta": np.array([*dims, *ws, *shift_cfg], dtype=np.int64),
           "shift": np.array(shift, dtype=np.int64), "pads": np.array(pads, dtype=np.int64)}
    if any(s > 0 for s in shift):
        m = ref.get_attn_mask(sp, ws, shift, pads)           # [1,P,N,N] float
        out["mask_bits"] = np.packbits(m[0].numpy().astype(np.uint8), axis=-1)
        out["mask_shape"] = np.array(m.shape, dtype=np.int64)
    # index map: feed a volume whose value IS its flat index (+1, 0 = padding)
    vol = torch.arange(1, dims[0] * dims[1] * dims[2] + 1, dtype=torch.float64).reshape(1, 1, *dims)
    v = torch.nn.functional.pad(vol, tuple(reversed(pads)))
    if any(s > 0 for s in shift):
        v = torch.roll(v, shifts=tuple(-s for s in shift), dims=(2, 3, 4))
    w_ = ref.window_partition(v, ws)                         # [1,P,1,wh,ww,wd]
    out["index_map"] = (w_.reshape(w_.shape[1], -1).numpy().astype(np.int64) - 1)
    return out


def run_pe_case(ref, seed=5):
    torch.manual_seed(seed)
    pe = ref.RelativePE(embed_dim=16, num_heads=3, max_abs_pos=(4, 4, 2), max_cap_dist=(4, 4, 2),
                        max_prompts=2, tokens_per_prompt=3, use_token_params=True).double()
    out = {"sd." + k: v.detach().numpy() for k, v in pe.state_dict().items()}
    out["bias_prompt"] = pe(4, 4, 2, 6).detach().numpy()
    out["bias_content"] = pe(4, 4, 2, 0).detach().numpy()
    return out


def run_pair_case(ref, seed=11):
    """ConsecutiveSwinBlocks with PatchMerging (both merge_last_dim variants, plus an odd-everywhere map whose
    PatchMerging pads every axis, down.py:24-28): output AND the gradients of x, both prompts and every parameter
    (incl. merge.norm / merge.reduction, down.py:21-53) for the loss sum(out * go), computed in float64."""
    res = {}
    for tag, mld, dims in (("mld1", True, (8, 8, 4)), ("mld0", False, (7, 8, 4)), ("odd1", True, (7, 5, 3))):
        torch.manual_seed(seed)
        pair = ref.ConsecutiveSwinBlocks(hidden_channels=12, num_heads=2, pos_bias_embed_dim=16,
                                         max_prompts=1, tokens_per_prompt=8, window_size=(4, 4, 2),
                                         use_token_params=True, down=True, merge_last_dim=mld)
        g = torch.Generator().manual_seed(seed + 3)
        x = _rand_like_ref_inputs(g, (1, 12, *dims))
        p0 = _rand_like_ref_inputs(g, (1, 8, 12), 0.5)
        p1 = _rand_like_ref_inputs(g, (1, 8, 12), 0.5)
        for k, v in pair.state_dict().items():
            res[f"{tag}.sd.{k}"] = v.detach().clone().numpy()
        res[f"{tag}.x"], res[f"{tag}.p0"], res[f"{tag}.p1"] = x.numpy(), p0.numpy(), p1.numpy()
        pair64 = pair.double()
        x64 = x.double().requires_grad_(True)
        p64 = [p0.double().requires_grad_(True), p1.double().requires_grad_(True)]
        y = pair64(x64, tuple(p64))
        go = _rand_like_ref_inputs(g, tuple(y.shape))
        (y * go.double()).sum().backward()
        res[f"{tag}.out"] = y.detach().float().numpy()
        res[f"{tag}.go"] = go.numpy()
        res[f"{tag}.grad.x"] = x64.grad.float().numpy()
        res[f"{tag}.grad.p0"] = p64[0].grad.float().numpy()
        res[f"{tag}.grad.p1"] = p64[1].grad.float().numpy()
        for n, prm in pair64.named_parameters():
            res[f"{tag}.grad.{n}"] = (prm.grad.float().numpy() if prm.grad is not None
                                      else np.zeros(prm.shape, np.float32))
    return res


# name, C, heads, b, p, n_q, n_k, bias shape, mask shape (None = argument omitted)
DENSE_CASES = [
    ("block_form", 12, 4, 2, 3, 40, 40, (1, 4, 40, 40), (1, 3, 1, 40, 40)),     # RelativePE.forward + get_attn_mask shapes
    ("bias_only",  12, 2, 1, 2, 16, 16, (2, 16, 16), None),
    ("mask_only",  24, 4, 2, 2, 16, 24, None, (2, 2, 1, 16, 24)),               # cross attention: n_k != n_q
    ("plain",      12, 4, 1, 2, 16, 16, None, None),
    ("full_rank",  12, 2, 2, 2, 8, 16, (2, 2, 2, 8, 16), (2, 1, 2, 8, 1)),      # per-sample bias, mask broadcast along keys
]


def run_dense_cases(ref, seed=23):
    """The reference WindowAttention called with its literal arguments (window_attention.py:35-61): dense pos_bias / mask
    tensors of several broadcast shapes, or None.  Output and the gradients of q, k, v, pos_bias and every parameter for
    the loss sum(out * go), float64."""
    res = {}
    for name, C, heads, b, p, nq, nk, bshape, mshape in DENSE_CASES:
        torch.manual_seed(seed)
        mod = ref.WindowAttention(C, heads)
        g = torch.Generator().manual_seed(seed + len(name))
        q = _rand_like_ref_inputs(g, (b, p, nq, C))
        k = _rand_like_ref_inputs(g, (b, p, nk, C))
        v = _rand_like_ref_inputs(g, (b, p, nk, C))
        bias = _rand_like_ref_inputs(g, bshape, 0.7) if bshape else None
        mask = (torch.rand(mshape, generator=g) > 0.35).float() if mshape else None
        go = _rand_like_ref_inputs(g, (b, p, nq, C))
        for kk, vv in mod.state_dict().items():
            res[f"{name}.sd.{kk}"] = vv.detach().clone().numpy()
        for kk, vv in (("q", q), ("k", k), ("v", v), ("bias", bias), ("mask", mask), ("go", go)):
            if vv is not None:
                res[f"{name}.{kk}"] = vv.numpy()
        m64 = mod.double()
        leaves = [t.double().requires_grad_(True) for t in (q, k, v)]
        b64 = bias.double().requires_grad_(True) if bias is not None else None
        y = m64(*leaves, pos_bias=b64, mask=None if mask is None else mask.double())
        (y * go.double()).sum().backward()
        res[f"{name}.out"] = y.detach().float().numpy()
        for kk, t in zip("qkv", leaves):
            res[f"{name}.grad.{kk}"] = t.grad.float().numpy()
        if b64 is not None:
            res[f"{name}.grad.bias"] = b64.grad.float().numpy()
        for n, prm in m64.named_parameters():
            res[f"{name}.grad.{n}"] = prm.grad.float().numpy()
    return res


def main():
    ref = ref_loader.load()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    if "--dense-only" in sys.argv:
        np.savez_compressed(os.path.join(GOLDEN_DIR, "attn_dense.npz"), **run_dense_cases(ref))
        print("wrote attn_dense")
        return
    if "--pair-only" in sys.argv:
        np.savez_compressed(os.path.join(GOLDEN_DIR, "pair_merge.npz"), **run_pair_case(ref))
        print("wrote pair_merge")
        return
    for i, case in enumerate(BLOCK_CASES):
        data = run_block_case(ref, case, seed=100 + i)
        np.savez_compressed(os.path.join(GOLDEN_DIR, case[0] + ".npz"), **data)
        print("wrote", case[0], {k: v.shape for k, v in data.items() if k in ("x", "out")})
    for case in GEOM_CASES:
        np.savez_compressed(os.path.join(GOLDEN_DIR, case[0] + ".npz"), **run_geom_case(ref, case))
        print("wrote", case[0])
    np.savez_compressed(os.path.join(GOLDEN_DIR, "pe_small.npz"), **run_pe_case(ref))
    np.savez_compressed(os.path.join(GOLDEN_DIR, "pair_merge.npz"), **run_pair_case(ref))
    np.savez_compressed(os.path.join(GOLDEN_DIR, "attn_dense.npz"), **run_dense_cases(ref))
    print("done")


if __name__ == "__main__":
    main()
