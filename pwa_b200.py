"""Import shim: the package directory is named `medical-image-segmentation-with-visual-prompts_b200`
(not a Python identifier), so `import pwa_b200` loads it through importlib and aliases it."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("medical-image-segmentation-with-visual-prompts_b200")
sys.modules[__name__] = _pkg
