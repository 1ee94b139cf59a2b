"""Import alias: `import afb200` == the package directory
`spatiotemporal-deepfake-detection-for-live-video-calls_b200/` (whose mandated
name is not a Python identifier).  Sub-modules are aliased too, so that
`from afb200.classifier import B200Engine` yields the very same class objects as
the package's own relative imports."""
import importlib
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
if _here not in sys.path:
    sys.path.insert(0, _here)
_name = "spatiotemporal-deepfake-detection-for-live-video-calls_b200"
_pkg = importlib.import_module(_name)
for _k, _v in list(sys.modules.items()):
    if _k.startswith(_name + "."):
        sys.modules["afb200" + _k[len(_name):]] = _v
sys.modules[__name__] = _pkg
