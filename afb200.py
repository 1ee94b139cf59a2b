"""Import alias: `import afb200` == the package directory
`spatiotemporal-deepfake-detection-for-live-video-calls_b200/` (whose mandated
name is not a Python identifier)."""
import importlib
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
if _here not in sys.path:
    sys.path.insert(0, _here)
_pkg = importlib.import_module("spatiotemporal-deepfake-detection-for-live-video-calls_b200")
sys.modules[__name__] = _pkg
