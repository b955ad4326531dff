"""Import alias: `import vlb200` resolves to the package directory `video-learning-tf_b200/`
(its name is not a valid Python identifier, so it is loaded through importlib)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("video-learning-tf_b200")
sys.modules[__name__] = _pkg
sys.modules.setdefault("vlb200", _pkg)
