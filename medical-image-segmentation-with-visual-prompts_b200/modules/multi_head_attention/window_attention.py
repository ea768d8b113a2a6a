"""Prompted multi-head window attention on the fused sm_100a kernels.

Drop-in for the reference's WindowAttention (multi_head_attention/window_attention.py:11-61): same
constructor, same parameters (`to_q/to_k/to_v` without bias, `proj` with bias), same ValueError for
an incompatible head count.  The reference materialises the [B,P,h,N',N'] logits and makes ~10
elementwise passes over them; here QK^T, bias, multiplicative shift mask, softmax and PV run inside
one kernel (csrc/attn_*.cu) and only the N content tokens are queries (the reference cuts the prompt
rows right after the residual, swin_block.py:222-225).
"""
from __future__ import annotations

from typing import NamedTuple, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import functional as PF


class BiasTables(NamedTuple):
    """Compact position bias (RelativePE.tables) + window shape."""
    th: torch.Tensor
    tw: torch.Tensor
    td: torch.Tensor
    tok: Optional[torch.Tensor]
    ws: tuple


class WindowAttention(nn.Module):
    def __init__(self, dim: int, num_heads: int, attn_drop: float = 0.0, proj_drop: float = 0.0):
        super().__init__()
        if dim % num_heads != 0:
            raise ValueError('WindowAttention: The dimension is not compatible with the number of heads!')
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.to_q = nn.Linear(dim, dim, bias=False)
        self.to_k = nn.Linear(dim, dim, bias=False)
        self.to_v = nn.Linear(dim, dim, bias=False)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self.impl = PF.IMPL_AUTO

    def project_prompts(self, prompts: Optional[torch.Tensor], lowp: Optional[dict] = None):
        """K|V projection of the normalised prompt tokens [B,I,C] -> [B,I,2C] (ONE [C -> 2C] GEMM; once per sample, the
        reference projects the prompt rows of every window)."""
        if prompts is None:
            return None
        return PF.multi_linear(prompts, None, self.to_k.weight, self.to_v.weight, lowp=(lowp or {}).get('kv'))

    def _drop_rates(self):
        return (float(self.attn_drop.p), float(self.proj_drop.p)) if self.training else (0.0, 0.0)

    def attend_packed(self, qkv: torch.Tensor, pos_bias: BiasTables, mask, prompt_kv, drop_seed=None):
        """The attention core alone: already projected q|k|v [B,P,N,3C] (+ prompt K|V [B,I,2C]) -> [B,P,N,C], attention
        dropout seeded by words 0-1 of `drop_seed`.  (The block runs this between its two checkpointed token segments.)"""
        p_drop, _ = self._drop_rates()
        if p_drop > 0 and drop_seed is None:
            drop_seed = PF.new_dropout_seed(qkv.device, 4)
        return PF.prompted_window_attention_packed(qkv, prompt_kv, pos_bias.th, pos_bias.tw, pos_bias.td, pos_bias.tok, mask,
                                                   self.num_heads, pos_bias.ws, self.scale, self.impl, p_drop=p_drop,
                                                   seed=None if drop_seed is None else drop_seed[:2])

    def project_out(self, o: torch.Tensor, lowp: dict, proj_bias_grad: bool = True, drop_seed=None, dropout: bool = True):
        """Output projection with its bias and (dropout=True) the projection dropout seeded by words 2-3 of `drop_seed`
        (reference :59-60).  dropout=False leaves the dropout to the caller's fused kernel."""
        _, p_proj = self._drop_rates()
        if not dropout or p_proj == 0:
            return PF.multi_linear(o, self.proj.bias, self.proj.weight, lowp=lowp.get('proj'),
                                   bias_grad=proj_bias_grad and dropout, lowp_bias=lowp.get('proj_b'))
        if drop_seed is None:
            drop_seed = PF.new_dropout_seed(o.device, 4)
        # with projection dropout the gradient of proj.bias is a by-product of the dropout's backward pass
        drop_db = proj_bias_grad and PF.dropout_colsum_supported(o.shape[-1])
        a = PF.multi_linear(o, self.proj.bias, self.proj.weight, lowp=lowp.get('proj'),
                            bias_grad=proj_bias_grad and not drop_db, lowp_bias=lowp.get('proj_b'))
        return PF.seeded_dropout(a, p_proj, drop_seed[2:4], bias_of_x=self.proj.bias if drop_db else None)

    def _forward_dense(self, q, k, v, pos_bias, mask, prompts, prompt_kv, drop_seed):
        if prompts is not None or prompt_kv is not None:
            raise ValueError("WindowAttention: with a dense pos_bias / mask the prompt rows are part of q, k, v "
                             "(reference window_attention.py:35-41); `prompts` belongs to the BiasTables form")
        if mask is not None and mask.dtype == torch.uint8 and mask.dim() == 2:
            raise ValueError("WindowAttention: uint8 region ids need the BiasTables form of pos_bias")
        p_drop = float(self.attn_drop.p) if self.training else 0.0
        p_proj = float(self.proj_drop.p) if self.training else 0.0
        if (p_drop > 0 or p_proj > 0) and drop_seed is None:
            drop_seed = PF.new_dropout_seed(q.device, 4)
        qq = PF.multi_linear(q, None, self.to_q.weight)
        kk = PF.multi_linear(k, None, self.to_k.weight)
        vv = PF.multi_linear(v, None, self.to_v.weight)
        o = PF.dense_window_attention(qq, kk, vv, pos_bias, mask, self.num_heads, self.scale, p_drop,
                                      None if drop_seed is None else drop_seed[:2])
        o = PF.multi_linear(o, self.proj.bias, self.proj.weight)
        return PF.seeded_dropout(o, p_proj, drop_seed[2:4]) if p_proj > 0 else o

    def forward(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, pos_bias=None,
                mask: Optional[torch.Tensor] = None, prompts: Optional[torch.Tensor] = None, lowp: Optional[dict] = None,
                proj_bias_grad: bool = True, drop_seed: Optional[torch.Tensor] = None,
                prompt_kv: Optional[torch.Tensor] = None):
        """Two argument forms.
        (1) The reference's (window_attention.py:35-41): q [b,p,n_q,C], k / v [b,p,n_k,C] (prompt rows included by the
        caller), `pos_bias` / `mask` dense tensors or None that broadcast against [b,p,h,n_q,n_k] (e.g. RelativePE.forward's
        [1,h,N',N'] and get_attn_mask's [1,P,1,N',N']); every row of q is a query.  Runs the dense-argument kernels
        (csrc/attn_dense.cu), differentiable in the bias as well.
        (2) The block's compact form, below: the fused kernels.
        q = k = v: normalised window tokens [B,P,N,C]; `prompts`: normalised prompt tokens [B,I,C]
        appended to the keys/values of every window; pos_bias: BiasTables; mask: uint8 region ids [P,N]
        (mask[p,i,j] = ids[p,i]==ids[p,j]) or None; lowp: optional {'qkv','kv','proj'} weights already cast to the
        compute dtype (SwinTransformerBlock packs them once per forward); drop_seed: optional int32 [4] device tensor:
        words 0-1 seed the attention dropout, words 2-3 the projection dropout (drawn here when absent); prompt_kv: `project_prompts(prompts, lowp)`
        computed by the caller ahead of time (self-attention path only), instead of `prompts`.  Returns [B,P,N,C]."""
        if not isinstance(pos_bias, BiasTables):
            # the reference's literal argument form (window_attention.py:35-58): dense tensors or None
            return self._forward_dense(q, k, v, pos_bias, mask, prompts, prompt_kv, drop_seed)
        # attention dropout (reference :57) happens INSIDE the fused kernel: the probabilities are never materialised.
        # The mask comes from a counter-based hash of (seed words, sample, window, head, query, key); it cannot be the
        # reference's torch Philox stream (that is indexed over a dense [B,P,h,N',N'] tensor which does not exist here).
        p_drop = float(self.attn_drop.p) if self.training else 0.0
        p_proj = float(self.proj_drop.p) if self.training else 0.0
        if (p_drop > 0 or p_proj > 0) and drop_seed is None:
            drop_seed = PF.new_dropout_seed(q.device, 4)
        proj_seed = None
        if drop_seed is not None:
            if drop_seed.numel() < 4:
                raise ValueError("WindowAttention: drop_seed must hold four int32 words (attention, projection)")
            drop_seed, proj_seed = drop_seed[:2], drop_seed[2:4]
        impl = self.impl
        if q is k and k is v:
            # self-attention (the only way the block calls it): ONE fused [C -> 3C] projection GEMM; the kernels
            # read q|k|v as column blocks of its output (row stride 3C), prompt K/V likewise from [C -> 2C]
            lp = lowp or {}
            qkv = PF.multi_linear(q, None, self.to_q.weight, self.to_k.weight, self.to_v.weight, lowp=lp.get('qkv'))
            kvp = prompt_kv if prompt_kv is not None else self.project_prompts(prompts, lp)
            o = PF.prompted_window_attention_packed(qkv, kvp, pos_bias.th, pos_bias.tw, pos_bias.td, pos_bias.tok, mask,
                                                    self.num_heads, pos_bias.ws, self.scale, impl, p_drop=p_drop, seed=drop_seed)
        else:
            qq = PF.multi_linear(q, None, self.to_q.weight)
            kk = PF.multi_linear(k, None, self.to_k.weight)
            vv = PF.multi_linear(v, None, self.to_v.weight)
            kp = vp = None
            if prompts is not None:
                kp = PF.multi_linear(prompts, None, self.to_k.weight)
                vp = PF.multi_linear(prompts, None, self.to_v.weight)
            o = PF.prompted_window_attention(qq, kk, vv, kp, vp, pos_bias.th, pos_bias.tw, pos_bias.td, pos_bias.tok,
                                             mask, self.num_heads, pos_bias.ws, self.scale, impl, p_drop=p_drop, seed=drop_seed)
        # with projection dropout the gradient of proj.bias is a by-product of the dropout's backward pass
        drop_db = p_proj > 0 and proj_bias_grad and PF.dropout_colsum_supported(o.shape[-1])
        o = PF.multi_linear(o, self.proj.bias, self.proj.weight, lowp=(lowp or {}).get('proj'),
                            bias_grad=proj_bias_grad and not drop_db, lowp_bias=(lowp or {}).get('proj_b'))
        # projection dropout (reference :60): seeded kernel instead of nn.Dropout, so that a checkpointed block inside a
        # CUDA graph recomputes the same mask without touching the generator state (csrc/dropout.cu)
        return PF.seeded_dropout(o, p_proj, proj_seed, bias_of_x=self.proj.bias if drop_db else None) if p_proj > 0 else o
