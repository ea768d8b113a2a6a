from .swin_transformer import (ConsecutiveSwinBlocks, SwinTransformerBlock, PatchMerging, window_partition,
                               window_reverse, get_attn_mask)
from .multi_head_attention import WindowAttention, RelativePE, BiasTables
from .swin_unetr import SwinUnetR, SwinUnetRConfig, SwinUpBlock
