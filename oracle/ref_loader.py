"""Import the LIVE reference hot-path modules without executing its package __init__.

Test infrastructure only.  `/root/reference/src/modules/__init__.py` eagerly imports
MONAI / torchinfo / matplotlib (absent here), so the hot-path files
(`swin_transformer/swin_block.py`, `multi_head_attention/*.py`, `swin_transformer/down.py`)
are loaded under a synthetic empty parent package whose `__path__` points at the
reference's `src/modules` directory.  Relative imports inside the reference
(`swin_block.py:8-9`) then resolve normally and no third-party code runs.

The reference tree does not exist on the GPU box: callers must check `available()`.
"""
import importlib
import os
import sys
import types

REF_MODULES_DIR = os.environ.get("PWA_REFERENCE_MODULES", "/root/reference/src/modules")
_PARENT = "_pwa_live_reference"


def available() -> bool:
    return os.path.isfile(os.path.join(REF_MODULES_DIR, "swin_transformer", "swin_block.py"))


def _ensure_parent():
    if _PARENT not in sys.modules:
        parent = types.ModuleType(_PARENT)
        parent.__path__ = [REF_MODULES_DIR]
        sys.modules[_PARENT] = parent


def load():
    """Returns a namespace with the reference's hot-path symbols."""
    if not available():
        raise RuntimeError(f"live reference not found under {REF_MODULES_DIR}")
    _ensure_parent()
    sb = importlib.import_module(f"{_PARENT}.swin_transformer.swin_block")
    wa = importlib.import_module(f"{_PARENT}.multi_head_attention.window_attention")
    pe = importlib.import_module(f"{_PARENT}.multi_head_attention.relative_positional_encoding")
    dn = importlib.import_module(f"{_PARENT}.swin_transformer.down")
    ns = types.SimpleNamespace(
        swin_block=sb,
        SwinTransformerBlock=sb.SwinTransformerBlock,
        ConsecutiveSwinBlocks=sb.ConsecutiveSwinBlocks,
        window_partition=sb.window_partition,
        window_reverse=sb.window_reverse,
        get_attn_mask=sb.get_attn_mask,
        WindowAttention=wa.WindowAttention,
        RelativePE=pe.RelativePE,
        PatchMerging=dn.PatchMerging,
    )
    return ns
