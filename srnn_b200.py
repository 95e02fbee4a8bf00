"""Import alias for the hyphenated package directory ``jalil-saboorizadeh-multi-speaker-neural-vocoder_b200``."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_pkg = importlib.import_module("jalil-saboorizadeh-multi-speaker-neural-vocoder_b200")
globals().update({k: getattr(_pkg, k) for k in _pkg.__all__})
package = _pkg
