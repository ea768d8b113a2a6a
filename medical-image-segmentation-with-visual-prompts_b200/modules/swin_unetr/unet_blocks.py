"""Decoder stage of the prompted SwinUNETR: the caller of the hot path on the way up.

Drop-in for the reference's `SwinUpBlock` (swin_unetr/unet_blocks.py:11-91) WITHOUT its MONAI imports: with the
reference's fixed arguments, `get_act_layer("leakyrelu")` is `nn.LeakyReLU()`, `get_norm_layer("batch", 3, ch)` is
`nn.BatchNorm3d(ch)` and `Convolution(..., conv_only=True)` is a `Sequential` whose only child is named `conv`
(unet_blocks.py:36-56), so the state-dict keys `norm_concat.*`, `conv_concat.conv.*`, `swin_layer.*` are kept.
Trilinear upsampling, concat, BatchNorm, LeakyReLU and the 3^3 convolution are plain library calls (cuDNN / ATen):
they are adjacent to, not on, the path this repo rebuilds; the prompted window-attention pair is
`ConsecutiveSwinBlocks(down=False)` on the sm_100a kernels.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Sequence

import torch
import torch.nn as nn

from ..swin_transformer import ConsecutiveSwinBlocks


class SwinUpBlock(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, strides: Sequence[int], kernel_size: Sequence[int],
                 pos_bias_embed_dim: int, num_heads: int, window_size: Sequence[int], max_prompts: int,
                 tokens_per_prompt: int, use_token_params: bool = True, act: str = "leakyrelu", norm: str = "batch",
                 attn_drop: float = 0.0, proj_drop: float = 0.0, use_checkpoint: bool = False, hidden_channels=None):
        super().__init__()
        if act != "leakyrelu" or norm != "batch":
            raise NotImplementedError("SwinUpBlock: only the reference's act='leakyrelu' / norm='batch' are built")
        self.up = nn.Upsample(scale_factor=tuple(strides), mode='trilinear', align_corners=False)
        self.act = nn.LeakyReLU()
        if hidden_channels is None:
            hidden_channels = in_channels + in_channels // 2
        self.norm_concat = nn.BatchNorm3d(hidden_channels)
        ks = tuple(kernel_size)
        self.conv_concat = nn.Sequential(OrderedDict(conv=nn.Conv3d(hidden_channels, out_channels, kernel_size=ks, stride=1,
                                                                    padding=tuple(k // 2 for k in ks))))
        self.swin_layer = ConsecutiveSwinBlocks(hidden_channels=out_channels, pos_bias_embed_dim=pos_bias_embed_dim,
                                                num_heads=num_heads, window_size=window_size, max_prompts=max_prompts,
                                                tokens_per_prompt=tokens_per_prompt, use_token_params=use_token_params,
                                                down=False, attn_drop=attn_drop, proj_drop=proj_drop,
                                                use_checkpoint=use_checkpoint)

    def forward(self, x, c, p=(None, None)):
        x = self.up(x)
        x = torch.cat([x[..., :c.size(2), :c.size(3), :c.size(4)], c], dim=1)      # crop to the skip's size (:73)
        x = self.conv_concat(self.act(self.norm_concat(x)))
        return self.swin_layer(x, p)

    def named_parameters_body(self):
        return [*self.norm_concat.named_parameters(), *self.conv_concat.named_parameters(),
                *self.swin_layer.named_parameters_body()]

    def named_parameters_bias_content(self):
        return self.swin_layer.named_parameters_bias_content()

    def named_parameters_bias_prompt_tokens(self):
        return self.swin_layer.named_parameters_bias_prompt_tokens()
