"""`from criterions import ctc, asg, stc, transducer` (reference utils.py:19) -> the product's criteria."""
import sys as _sys

from gtn_applications_b200.criterions import asg, ctc, stc, transducer  # noqa: F401

for _n, _m in (("ctc", ctc), ("asg", asg), ("stc", stc), ("transducer", transducer)):
    _sys.modules[__name__ + "." + _n] = _m
