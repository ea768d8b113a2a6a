"""Import shim: the package directory is named `medical-image-segmentation-with-visual-prompts_b200`
(not a Python identifier), so `import pwa_b200` loads it through importlib and aliases it -- including every
submodule, so that `pwa_b200.functional` and `<package>.functional` are the SAME module object."""
import importlib
import os
import sys

_REAL = "medical-image-segmentation-with-visual-prompts_b200"
_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module(_REAL)
for _name, _mod in list(sys.modules.items()):
    if _name == _REAL or _name.startswith(_REAL + "."):
        sys.modules[__name__ + _name[len(_REAL):]] = _mod
sys.modules[__name__] = _pkg
