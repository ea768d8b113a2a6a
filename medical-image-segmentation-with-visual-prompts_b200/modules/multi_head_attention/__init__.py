from .window_attention import WindowAttention, BiasTables
from .relative_positional_encoding import RelativePE

__all__ = ['WindowAttention', 'RelativePE', 'BiasTables']
