"""ORACLE — TEST INFRASTRUCTURE ONLY.  Float64 ("truth") build of the ``gtn``
shim in ``oracle/gtn``: same API, every weight / score / gradient in double.
Used to judge both the float32 oracle and the CUDA path."""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_root = os.path.dirname(_here)
if _root not in sys.path:
    sys.path.insert(0, _root)

import _gtn_oracle as _ext  # noqa: E402
from gtn._api import bind as _bind  # noqa: E402

_bind(globals(), _ext.f64)
