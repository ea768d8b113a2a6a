"""Backbone host around the hot path: patch embedding, three prompted Swin encoder stages, bottleneck, prompted Swin
decoder stages, SSL / segmentation heads and the prompt-token ParameterLists.

MONAI-free drop-in for the reference's `SwinUnetR` (swin_unetr/swin_unetr.py:8-527): same constructor (`conf`: any
object with the reference's config attributes, e.g. the Namespace `utils/configs.py:13` builds or `SwinUnetRConfig`
below), same submodule / parameter names (state-dict compatible: `input_layer.{0,1}`, `encoder_blocks.N`, `bottleneck`,
`residual_blocks.N`, `decoder_blocks.N`, `output_layer`, `prompt_tokens.{enc,dec,out}.N`, `extra_heads.*`), same
forward outputs per training mode (:129-144), same freeze logic (:21-40) and parameter-group accessors (:434-527) the
trainers' optimisers use.  Everything that is not the prompted window-attention pair is a plain library call
(Conv3d / BatchNorm3d / InstanceNorm3d / Upsample: SURVEY.md §8f-3, adjacent to the path).

Built: the reference's default wiring `unetr_res_block: none|simple`, `unetr_up_block: swin`
(configurations/example_configs.yml:8-9).  `unetr_res_block: full` and the CNN up blocks are MONAI `UnetrBasicBlock` /
`UnetrUpBlock` networks (:249-290, :339-383) and raise NotImplementedError.

Differences in HOW (not what): prompt tokens reach the blocks as broadcast views `[1,I,C] -> [B,I,C]` (`expand`)
instead of `.repeat` copies (:56-60, :93-96) -- autograd sums their gradient over the batch either way.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence

import torch
import torch.nn as nn

from ..swin_transformer import ConsecutiveSwinBlocks
from .unet_blocks import SwinUpBlock

_DECODER_MODES = ('self_supervised_learning_decoder', 'supervised_learning_decoder')
_ALL_MODES = ('self_supervised_learning_all', 'supervised_learning_all')
_SUPERVISED = ('supervised_learning_decoder', 'supervised_learning_all')


@dataclass
class SwinUnetRConfig:
    """The config attributes `SwinUnetR` reads, with the values of configurations/example_configs.yml:1-24,64-72,98."""
    training_mode: str = 'self_supervised_learning_encoder'
    input_channels: int = 1
    depth_unet: int = 3
    hidden_channels: List[int] = field(default_factory=lambda: [48, 96, 192, 384])
    input_patch_size: Sequence[int] = (2, 2, 2)
    unetr_res_block: str = 'none'
    unetr_up_block: str = 'swin'
    basic_block_res: bool = True
    num_heads_encoder: int = 4
    num_heads_decoder: int = 4
    attn_window_size: Sequence[int] = (8, 8, 4)
    pos_bias_embed_dim: int = 64
    use_checkpoint: bool = True
    attn_drop: float = 0.1
    proj_drop: float = 0.1
    max_prompts: int = 1
    tokens_per_prompt_encoder: int = 64
    tokens_per_prompt_decoder: int = 64
    use_encoder_prompting: bool = False
    use_decoder_prompting: bool = False
    contrastive_coding_dim: int = 512
    use_reconstruction: bool = True
    use_rotation_prediction: bool = True
    use_contrastive_learning: bool = True
    use_mutual_learning: bool = False
    output_channels_pretrain: int = 5
    output_channels_downstream: int = 2


def _xavier_tokens(n_tokens, channels):
    return nn.Parameter(nn.init.xavier_uniform_(torch.empty((n_tokens, channels)), gain=nn.init.calculate_gain('linear')),
                        requires_grad=True)


def _broadcast(tokens, batch):
    """[I,C] parameter -> [B,I,C] view (the reference materialises B copies with .repeat)."""
    return tokens.unsqueeze(0).expand(batch, -1, -1)


class SwinUnetR(nn.Module):
    def __init__(self, conf):
        super().__init__()
        self.input_layer = None
        self.encoder_blocks = None
        self.bottleneck = None
        self.residual_blocks = None
        self.decoder_blocks = None
        self.output_layer = None
        self.prompt_tokens = nn.ModuleDict()
        self.extra_heads = nn.ModuleDict()
        self.conf = conf
        mode = conf.training_mode
        if mode == 'self_supervised_learning_encoder':
            self.setup_ssl_encoder()
        elif mode in _DECODER_MODES:
            self.setup_ssl_decoder()
            for _, p in self.named_parameters_encoder(include_prompt_tokens=conf.use_encoder_prompting):
                p.requires_grad = False
        elif mode in _ALL_MODES:
            self.setup_ssl_decoder()
        elif mode == 'downstream':
            # frozen backbone: only prompt tokens, their bias parameters and the head train (:33-40, :434-441)
            self.setup_downstream()
            for _, p in self.named_parameters_encoder(include_prompt_tokens=False):
                p.requires_grad = False
            for _, p in self.named_parameters_decoder(include_prompt_tokens=False):
                p.requires_grad = False
        else:
            raise ValueError(f'Training mode {mode} not available!')

    # ------------------------------------------------------------------------------------------------
    # forward paths
    # ------------------------------------------------------------------------------------------------
    def _prompts(self, group, j, batch, enabled):
        if not enabled:
            return [None, None]
        toks = self.prompt_tokens[group]
        return [_broadcast(toks[2 * j], batch), _broadcast(toks[2 * j + 1], batch)]

    def forward_swin_transformer(self, x):
        """out_vit = [deepest stage output, ..., stage-0 output, patch embedding, x] (reference :46-63)."""
        outs = [x]
        # what the stages need besides the feature map (bias tables, packed weights, prompt K|V) depends on parameters only:
        # under graph capture all stages start computing it on the side stream now, under the patch embedding
        ps = [self._prompts('enc', j, x.size(0), self.conf.use_encoder_prompting) for j in range(self.conf.depth_unet)]
        for j in range(self.conf.depth_unet):
            self.encoder_blocks[j].prefetch_side_inputs(ps[j], x)
        enc = self.input_layer(x)
        outs.insert(0, enc)
        for j in range(self.conf.depth_unet):
            enc = self.encoder_blocks[j](enc, ps[j])
            outs.insert(0, enc)
        return {'out_vit': outs}

    def forward_ssl_encoder(self, x):
        out = {}
        out_vit = self.forward_swin_transformer(x)['out_vit']
        c = self.conf
        if c.training_mode == 'self_supervised_learning_encoder':
            if c.use_reconstruction or c.use_mutual_learning:
                out['reconstruction'] = self.extra_heads['reconstruction'](out_vit[0])
            if c.use_rotation_prediction or c.use_contrastive_learning:
                pooled = out_vit[0].mean(dim=(2, 3, 4))                    # AdaptiveAvgPool3d((1,1,1)) + squeezes (:74-81)
                if c.use_rotation_prediction:
                    out['rotation_prediction'] = self.extra_heads['rotation_prediction'](pooled)
                if c.use_contrastive_learning:
                    out['contrastive_coding'] = self.extra_heads['contrastive_coding'](pooled)
        out['out_vit'] = out_vit
        return out

    def forward_decoder(self, c):
        conf = self.conf
        # (as in forward_swin_transformer: the decoder stages' feature-map-independent inputs start on the side stream now)
        ps = [self._prompts('dec', j, c[0].size(0), conf.use_decoder_prompting) for j in range(conf.depth_unet)]
        for j in range(conf.depth_unet):
            swin = getattr(self.decoder_blocks[j], 'swin_layer', None)
            if swin is not None:
                swin.prefetch_side_inputs(ps[j], c[0])
        dec = self.bottleneck(c[0]) + c[0]
        for j in range(conf.depth_unet):
            res = self.residual_blocks[j](c[j + 1])
            dec = self.decoder_blocks[j](dec, res, ps[j])
        if conf.unetr_res_block == 'none':
            out = self.output_layer(dec)
        else:
            out = self.output_layer(dec, self.residual_blocks[-1](c[-1]),
                                    self._prompts('out', 0, dec.size(0), conf.use_decoder_prompting))
        return {'latent_outputs': out}

    def forward_ssl_decoder(self, x):
        out_dec = self.forward_decoder(self.forward_ssl_encoder(x)['out_vit'])
        if self.conf.training_mode in _SUPERVISED:
            out_dec['seg_pred'] = self.extra_heads['segmentation'](out_dec['latent_outputs'])
        return out_dec

    def forward_downstream(self, x):
        return {'downstream': self.extra_heads['downstream'](self.forward_ssl_decoder(x)['latent_outputs'])}

    def forward(self, x):
        mode = self.conf.training_mode
        if mode == 'self_supervised_learning_encoder':
            return self.forward_ssl_encoder(x)
        if mode in _DECODER_MODES or mode in _ALL_MODES:
            return self.forward_ssl_decoder(x)
        if mode == 'downstream':
            return self.forward_downstream(x)
        raise ValueError(f'Training mode {mode} not available!')

    # ------------------------------------------------------------------------------------------------
    # construction
    # ------------------------------------------------------------------------------------------------
    def setup_swin_transformer(self, in_chs):
        c = self.conf
        self.input_layer = nn.Sequential(
            nn.Conv3d(c.input_channels, c.hidden_channels[0], kernel_size=tuple(c.input_patch_size),
                      stride=tuple(c.input_patch_size)),
            nn.BatchNorm3d(c.hidden_channels[0], eps=1e-6))
        # only the first stage halves the last axis as well (:160-161)
        self.encoder_blocks = nn.ModuleList([
            ConsecutiveSwinBlocks(hidden_channels=in_chs[i], pos_bias_embed_dim=c.pos_bias_embed_dim,
                                  num_heads=c.num_heads_encoder * (2 ** i), window_size=c.attn_window_size,
                                  max_prompts=c.max_prompts, tokens_per_prompt=c.tokens_per_prompt_encoder,
                                  use_token_params=c.use_encoder_prompting, down=True, merge_last_dim=(i < 1),
                                  attn_drop=c.attn_drop, proj_drop=c.proj_drop, use_checkpoint=c.use_checkpoint)
            for i in range(c.depth_unet)])

    def setup_ssl_encoder(self):
        c = self.conf
        in_chs = [c.hidden_channels[i] for i in range(c.depth_unet)]
        self.setup_swin_transformer(in_chs)
        n = len(in_chs)
        if c.use_reconstruction or c.use_mutual_learning:
            # conv -> InstanceNorm -> LeakyReLU -> x2 upsample, n + 1 times (the last one also along depth), 1x1 conv (:187-211)
            chs = [c.hidden_channels[-1] // (2 ** i) for i in range(n + 1)] + [c.hidden_channels[-1] // (2 ** n)]
            layers = []
            for i in range(n + 1):
                layers += [nn.Conv3d(chs[i], chs[i + 1], kernel_size=3, stride=1, padding=1), nn.InstanceNorm3d(chs[i + 1]),
                           nn.LeakyReLU(),
                           nn.Upsample(scale_factor=(2, 2, 1 if i < n - 1 else 2), mode='trilinear', align_corners=True)]
            layers.append(nn.Conv3d(chs[-1], c.input_channels, kernel_size=1, stride=1))
            self.extra_heads['reconstruction'] = nn.Sequential(*layers)
        if c.use_rotation_prediction:
            self.extra_heads['rotation_prediction'] = nn.Linear(c.hidden_channels[-1], 4)
        if c.use_contrastive_learning:
            self.extra_heads['contrastive_coding'] = nn.Linear(c.hidden_channels[-1], c.contrastive_coding_dim)
        if c.use_encoder_prompting:
            self.setup_prompt_tokens_encoder()

    def setup_downstream(self):
        c = self.conf
        self.setup_ssl_decoder()
        self.extra_heads['downstream'] = nn.Sequential(
            nn.BatchNorm3d(c.hidden_channels[0]),
            nn.Conv3d(c.hidden_channels[0], c.output_channels_downstream, kernel_size=3, stride=1, padding=1))

    def setup_ssl_decoder(self):
        c = self.conf
        if c.unetr_res_block == 'full' or c.unetr_up_block != 'swin':
            raise NotImplementedError("SwinUnetR: unetr_res_block='full' and CNN up blocks are MONAI UnetrBasicBlock / "
                                      "UnetrUpBlock networks (reference :249-383); built here: unetr_res_block in "
                                      "{'none','simple'} with unetr_up_block='swin' (the reference's example config)")
        in_chs = [c.hidden_channels[i] for i in range(c.depth_unet)]
        out_chs = [c.hidden_channels[i + 1] for i in range(c.depth_unet)]
        self.setup_swin_transformer(in_chs)
        in_chs.reverse()
        out_chs.reverse()
        self.bottleneck = nn.Conv3d(out_chs[0], out_chs[0], kernel_size=3, stride=1, padding=1)
        if c.unetr_res_block == 'simple':
            self.residual_blocks = nn.ModuleList(
                [nn.Conv3d(in_chs[i], in_chs[i], kernel_size=3, stride=1, padding=1) for i in range(c.depth_unet)]
                + [nn.Conv3d(c.input_channels, in_chs[-1], kernel_size=3, stride=1, padding=1)])
        else:
            self.residual_blocks = nn.ModuleList([nn.Identity() for _ in range(c.depth_unet + 1)])
        up = dict(kernel_size=(3, 3, 3), pos_bias_embed_dim=c.pos_bias_embed_dim, num_heads=c.num_heads_decoder,
                  window_size=c.attn_window_size, max_prompts=c.max_prompts, tokens_per_prompt=c.tokens_per_prompt_decoder,
                  attn_drop=c.attn_drop, proj_drop=c.proj_drop, use_checkpoint=c.use_checkpoint)
        # the last decoder stage undoes the first encoder stage's depth halving (:314-315)
        self.decoder_blocks = nn.ModuleList([
            SwinUpBlock(in_channels=out_chs[i], out_channels=in_chs[i], strides=(2, 2, 1 if i < len(in_chs) - 1 else 2),
                        use_token_params=c.use_decoder_prompting, **up)
            for i in range(c.depth_unet)])
        if c.unetr_res_block == 'none':
            self.output_layer = nn.Upsample(scale_factor=(2, 2, 2), mode='trilinear', align_corners=False)
        else:
            self.output_layer = SwinUpBlock(in_channels=in_chs[-1], out_channels=in_chs[-1], hidden_channels=2 * in_chs[-1],
                                            strides=(2, 2, 2), **up)
        if c.training_mode in _SUPERVISED:
            self.extra_heads['segmentation'] = nn.Sequential(
                nn.BatchNorm3d(c.hidden_channels[0]),
                nn.Conv3d(c.hidden_channels[0], c.output_channels_pretrain, kernel_size=3, stride=1, padding=1))
        if c.use_encoder_prompting:
            self.setup_prompt_tokens_encoder()
        if c.use_decoder_prompting:
            self.setup_prompt_tokens_decoder()

    def setup_prompt_tokens_encoder(self):
        c = self.conf
        if c.use_encoder_prompting:                  # two token sets per stage: unshifted / shifted block (:400-409)
            self.prompt_tokens['enc'] = nn.ParameterList(
                [_xavier_tokens(c.tokens_per_prompt_encoder, c.hidden_channels[i // 2]) for i in range(2 * c.depth_unet)])

    def setup_prompt_tokens_decoder(self):
        c = self.conf
        if c.use_decoder_prompting:
            self.prompt_tokens['dec'] = nn.ParameterList(
                [_xavier_tokens(c.tokens_per_prompt_decoder, c.hidden_channels[-(i + 1) // 2 - 1]) for i in range(2 * c.depth_unet)])
            if c.unetr_res_block != 'none' and c.unetr_up_block == 'swin':
                self.prompt_tokens['out'] = nn.ParameterList(
                    [_xavier_tokens(c.tokens_per_prompt_decoder, c.hidden_channels[0]) for _ in range(2)])

    # ------------------------------------------------------------------------------------------------
    # parameter groups (reference :434-527; consumed by the trainers' optimisers and the freeze logic above)
    # ------------------------------------------------------------------------------------------------
    def named_parameters_downstream(self):
        params = []
        if self.conf.use_encoder_prompting:
            params.extend(self.named_parameters_prompt_tokens_encoder())
        if self.conf.use_decoder_prompting:
            params.extend(self.named_parameters_prompt_tokens_decoder())
        params.extend(self.extra_heads['downstream'].named_parameters())
        return params

    def named_parameters_prompt_tokens_encoder(self):
        out = list(self.prompt_tokens['enc'].named_parameters())
        for blk in self.encoder_blocks:
            out.extend(blk.named_parameters_bias_prompt_tokens())
        return out

    def named_parameters_prompt_tokens_decoder(self):
        c = self.conf
        tokens = list(self.prompt_tokens['dec'].named_parameters())
        bias = []
        for blk in self.decoder_blocks:
            bias.extend(blk.named_parameters_bias_prompt_tokens())
        if c.unetr_res_block != 'none' and c.unetr_up_block == 'swin':
            tokens.extend(self.prompt_tokens['out'].named_parameters())
        if c.unetr_res_block != 'none':
            bias.extend(self.output_layer.named_parameters_bias_prompt_tokens())
        return [*tokens, *bias]

    def named_parameters_encoder(self, include_prompt_tokens=False):
        params = list(self.input_layer.named_parameters())
        for blk in self.encoder_blocks:
            params.extend(blk.named_parameters_body())
            params.extend(blk.named_parameters_bias_content())
        if include_prompt_tokens and self.conf.use_encoder_prompting:
            params.extend(self.named_parameters_prompt_tokens_encoder())
        if self.conf.training_mode == 'self_supervised_learning_encoder':
            for head in self.extra_heads.values():
                params.extend(head.named_parameters())
        return params

    def named_parameters_decoder(self, include_prompt_tokens=False):
        c = self.conf
        params = list(self.bottleneck.named_parameters())
        for blk in self.residual_blocks:
            params.extend(blk.named_parameters())
        for blk in self.decoder_blocks:
            params.extend(blk.named_parameters_body())
            params.extend(blk.named_parameters_bias_content())
        if c.unetr_res_block != 'none':
            params.extend(self.output_layer.named_parameters_body())
            params.extend(self.output_layer.named_parameters_bias_content())
        if include_prompt_tokens and c.use_decoder_prompting:
            params.extend(self.named_parameters_prompt_tokens_decoder())
        if c.training_mode in _SUPERVISED:
            params.extend(self.extra_heads['segmentation'].named_parameters())
        return params
