"""ORACLE — TEST INFRASTRUCTURE ONLY.

A Python package named ``gtn`` that exposes the CPU restatement of the GTN
ops (``oracle/gtn_cpu.h``) under the names the reference imports
(``import gtn`` at criterions/ctc.py:9, asg.py:9, stc.py:8, transducer.py:8,
utils.py:13).  With ``oracle/`` on ``sys.path`` the reference's
``criterions/*.py`` and ``tests/*.py`` import UNCHANGED against it; that is how
the oracle is pinned to the reference's golden values
(tests/test_oracle_reference.py).

Nothing under ``gtn_applications_b200/`` imports this package.  The scalar
type is float32, GTN's own; ``oracle/gtn64`` is the float64 truth build.
"""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_root = os.path.dirname(_here)
if _root not in sys.path:
    sys.path.insert(0, _root)

try:
    import _gtn_oracle as _ext
except ImportError as e:  # pragma: no cover
    raise ImportError(
        "oracle extension not built: run `make -C oracle` "
        "(or `python -c 'import __graft_entry__ as g; g.build()'`)"
    ) from e

from ._api import bind as _bind

_bind(globals(), _ext.f32)
