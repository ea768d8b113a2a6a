from .swin_block import (ConsecutiveSwinBlocks, SwinTransformerBlock, window_partition, window_reverse,
                         get_attn_mask)
from .down import PatchMerging

__all__ = ['ConsecutiveSwinBlocks', 'SwinTransformerBlock', 'PatchMerging', 'window_partition', 'window_reverse',
           'get_attn_mask']
