"""PatchMerging (reference swin_transformer/down.py:5-59): 2x2x2 (or 2x2x1) strided gather ->
LayerNorm(8C / 4C) -> Linear(no bias).  Adjacent to the hot path; kept API- and
state-dict-compatible so ConsecutiveSwinBlocks(down=True) chains stages.  Gather order of the
neighbourhood offsets (dh,dw,dd): 000,100,010,001,110,101,011,111 (:31-39); 00,10,01,11 when the last
axis is not merged (:41-45).  Odd axes get one zero plane on the LOW side (the reference reverses the
flat padding list before F.pad, :26-28, which swaps each (lo,hi) pair)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

_OFFS3 = ((0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (1, 0, 1), (0, 1, 1), (1, 1, 1))
_OFFS2 = ((0, 0), (1, 0), (0, 1), (1, 1))


class PatchMerging(nn.Module):
    def __init__(self, in_channels, out_channels, merge_last_dim=True):
        super().__init__()
        k = 8 if merge_last_dim else 4
        self.norm = nn.LayerNorm(k * in_channels, eps=1e-6)
        self.reduction = nn.Linear(k * in_channels, out_channels, bias=False)
        self.merge_last_dim = merge_last_dim

    def forward(self, x):
        h, w, d = x.shape[2:]
        if h % 2 or w % 2 or d % 2:
            x = F.pad(x, (d % 2, 0, w % 2, 0, h % 2, 0))
        if self.merge_last_dim:
            parts = [x[:, :, a::2, b::2, c::2] for a, b, c in _OFFS3]
        else:
            parts = [x[:, :, a::2, b::2, :] for a, b in _OFFS2]
        t = torch.cat(parts, dim=1).permute(0, 2, 3, 4, 1)             # [B,h/2,w/2,d',kC] channels last
        dt = t.dtype
        from ... import functional as PF
        if t.is_cuda and PF.layer_norm_supported(t.shape[-1]):
            t = PF.layer_norm(t.contiguous(), self.norm.weight, self.norm.bias, self.norm.eps)
            t = PF.multi_linear(t, None, self.reduction.weight)
        else:
            t = F.layer_norm(t, self.norm.normalized_shape, self.norm.weight.to(dt), self.norm.bias.to(dt), self.norm.eps)
            t = F.linear(t, self.reduction.weight.to(dt))
        return t.permute(0, 4, 1, 2, 3).contiguous()

    def named_parameters_body(self):
        return [*self.reduction.named_parameters(), *self.norm.named_parameters()]
