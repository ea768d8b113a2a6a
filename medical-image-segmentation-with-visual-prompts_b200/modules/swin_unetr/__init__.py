from .swin_unetr import SwinUnetR, SwinUnetRConfig
from .unet_blocks import SwinUpBlock

__all__ = ['SwinUnetR', 'SwinUnetRConfig', 'SwinUpBlock']
