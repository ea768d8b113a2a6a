"""Parameters for the CPU reference arm of bench.py (test / measurement infrastructure only): a state dict with the
reference's key names and shapes for one ConsecutiveSwinBlocks pair, initialised with plain torch initialisers the way the
reference does (xavier_uniform, gain 1, for every relative-position parameter, relative_positional_encoding.py:21-97;
torch defaults for Linear / LayerNorm, swin_block.py:128-143, down.py:11-19).  Nothing of the product is imported: the
reference arm must not map the repo's CUDA library."""
from __future__ import annotations

import math

import torch
import torch.nn as nn


def _xavier(*shape):
    return nn.init.xavier_uniform_(torch.empty(shape), gain=1.0)


def _linear(cout, cin, bias=True, prefix=""):
    lin = nn.Linear(cin, cout, bias=bias)
    out = {prefix + "weight": lin.weight.detach().clone()}
    if bias:
        out[prefix + "bias"] = lin.bias.detach().clone()
    return out


def block_state_dict(C, heads, E, I, ws):
    sd = {}
    for a, ax in enumerate("hwd"):
        sd[f"pe.enc_content_{ax}"] = _xavier(2 * ws[a] - 1, E)
    for ax in "hwd":
        sd[f"pe.weights_content_{ax}"] = _xavier(heads, E)
    if I > 0:
        sd["pe.enc_token.0"] = _xavier(I, E)
        sd["pe.weights_token"] = _xavier(heads, E)
    for name in ("attn_norm", "mlp_norm"):
        sd[f"{name}.weight"], sd[f"{name}.bias"] = torch.ones(C), torch.zeros(C)
    for name in ("to_q", "to_k", "to_v"):
        sd.update(_linear(C, C, bias=False, prefix=f"attn.{name}."))
    sd.update(_linear(C, C, prefix="attn.proj."))
    sd.update(_linear(C, C, prefix="mlp."))
    return sd


def pair_state_dict(C, heads, E, I, ws, down=True, merge_last_dim=True):
    sd = {}
    for i in range(2):
        sd.update({f"swin_blocks.{i}.{k}": v for k, v in block_state_dict(C, heads, E, I, ws).items()})
    if down:
        k = 8 if merge_last_dim else 4
        sd["merge.norm.weight"], sd["merge.norm.bias"] = torch.ones(k * C), torch.zeros(k * C)
        sd.update(_linear(2 * C, k * C, bias=False, prefix="merge.reduction."))
    return sd
