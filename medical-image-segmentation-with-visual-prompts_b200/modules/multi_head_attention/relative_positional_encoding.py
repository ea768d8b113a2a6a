"""Learned separable 3-axis relative position bias + per-prompt-token bias.

Drop-in for the reference's RelativePE (multi_head_attention/relative_positional_encoding.py:8-154):
same constructor, parameter / buffer names and shapes, initialisation and parameter-group accessors,
so reference checkpoints load unchanged.  The reference materialises a dense [1,h,N+I,N+I] tensor on
every forward (:99-142); the fused kernels instead consume three per-axis tables [h,w,w] and a [h,I]
prompt-column vector (`tables()`), from which bias[n][m] = th[ih][jh] + tw[iw][jw] + td[id][jd].
`forward()` still returns the dense form for API compatibility.
"""
from __future__ import annotations

from typing import Sequence

import torch
import torch.nn as nn

_AXES = "hwd"


def _xavier(*shape):
    return nn.Parameter(nn.init.xavier_uniform_(torch.empty(shape), gain=nn.init.calculate_gain('linear')))


class RelativePE(nn.Module):
    def __init__(self, embed_dim: int, num_heads: int, max_abs_pos: Sequence[int], max_cap_dist: Sequence[int],
                 max_prompts: int, tokens_per_prompt: int, use_token_params: bool = True):
        super().__init__()
        self.scale = embed_dim ** -0.5
        self.num_heads = num_heads
        # creation order follows the reference so that seeded initialisation matches bit for bit
        for a, ax in enumerate(_AXES):
            setattr(self, f"enc_content_{ax}", _xavier(2 * max_cap_dist[a] - 1, embed_dim))
        for a, ax in enumerate(_AXES):
            pos = torch.arange(max_abs_pos[a], dtype=torch.long)
            rel = (pos[None, :] - pos[:, None] + max_cap_dist[a] - 1).clamp_(0, 2 * (max_cap_dist[a] - 1))
            self.register_buffer(f"relative_dist_{ax}", rel)
        for ax in _AXES:
            setattr(self, f"weights_content_{ax}", _xavier(num_heads, embed_dim))
        if use_token_params:
            self.enc_token = nn.ParameterList([_xavier(tokens_per_prompt, embed_dim) for _ in range(max_prompts)])
            self.weights_token = _xavier(num_heads, embed_dim)

    # -- compact form used by the fused kernels ---------------------------------------------------
    def tables(self, dim_h: int, dim_w: int, dim_d: int, dim_i: int = 0):
        """(th [h,dim_h,dim_h], tw, td, tok [h,dim_i] | None), fp32, including the /3 and
        embed_dim**-0.5 factors (reference :116-123, :136-138)."""
        if self.enc_content_h.is_cuda:
            # one pwa kernel launch (csrc/bias.cu) instead of ~10 tiny torch ops (and ~25 in backward)
            from ... import functional as PF
            enc_tok = w_tok = None
            if dim_i > 0:
                enc_tok = self.enc_token[0] if len(self.enc_token) == 1 else torch.cat(list(self.enc_token), dim=0)
                if enc_tok.shape[0] != dim_i:
                    raise RuntimeError(f"RelativePE: {dim_i} prompt tokens given but max_prompts*tokens_per_prompt = "
                                       f"{enc_tok.shape[0]}")
                w_tok = self.weights_token
            dims = (dim_h, dim_w, dim_d)
            if not all(getattr(self, f"relative_dist_{ax}").shape[0] >= n for ax, n in zip(_AXES, dims)):
                raise RuntimeError(f"RelativePE: window {dims} exceeds max_abs_pos of the module")
            return PF.bias_tables(self.enc_content_h, self.enc_content_w, self.enc_content_d, self.weights_content_h,
                                  self.weights_content_w, self.weights_content_d, enc_tok, w_tok, dims)
        # Parameters in HOST memory (state-dict inspection, float64 checks of the table algebra against the reference in
        # tests/test_host_cpu.py): the same few-KB parameter -> table algebra in torch.  Not a compute path: the kernels
        # only ever see tables built on the device above, and the blocks raise on CPU tensors.
        out = []
        for ax, n in zip(_AXES, (dim_h, dim_w, dim_d)):
            enc = getattr(self, f"enc_content_{ax}")
            rel = getattr(self, f"relative_dist_{ax}")[:n, :n]
            w = getattr(self, f"weights_content_{ax}")
            # R[h,i,j] = sum_c w[h,c] * enc[rel[i,j], c]   -- one small GEMM over the 2w-1 distinct rows
            per_dist = (w.float() @ enc.float().t()) * (self.scale / 3.0)        # [h, 2w-1]
            out.append(per_dist[:, rel])                                          # [h, n, n]
        tok = None
        if dim_i > 0:
            enc_tok = torch.cat(list(self.enc_token), dim=0)
            if enc_tok.shape[0] != dim_i:
                raise RuntimeError(f"RelativePE: {dim_i} prompt tokens given but max_prompts*tokens_per_prompt = "
                                   f"{enc_tok.shape[0]}")
            tok = (self.weights_token.float() @ enc_tok.float().t()) * self.scale  # [h, I]
        return out[0], out[1], out[2], tok

    # -- dense form, reference signature -----------------------------------------------------------
    def forward(self, dim_h, dim_w, dim_d, dim_i=0):
        th, tw, td, tok = self.tables(dim_h, dim_w, dim_d, dim_i)
        n = dim_h * dim_w * dim_d
        content = (th[:, :, None, None, :, None, None] + tw[:, None, :, None, None, :, None]
                   + td[:, None, None, :, None, None, :]).reshape(1, self.num_heads, n, n)
        if dim_i == 0:
            return content
        total = content.new_zeros((1, self.num_heads, n + dim_i, n + dim_i))
        total[:, :, :n, :n] = content
        total[:, :, :n, n:] = tok[None, :, None, :]
        return total

    def named_parameters_bias_content(self):
        return [(n, p) for n, p in self.named_parameters() if 'enc_content' in n or 'weights_content' in n]

    def named_parameters_bias_prompt_tokens(self):
        return [(n, p) for n, p in self.named_parameters() if 'enc_token' in n or 'weights_token' in n]
