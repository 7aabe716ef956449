"""Imports the reference's criterions/*.py UNCHANGED from /root/reference against
the oracle `gtn` shim, with the compat names its bit-rotted tests / benchmarks
expect.  Build-container only."""
import importlib.util
import os
import sys

REFERENCE = "/root/reference"


def load_reference():
    """returns (ctc, asg, stc, transducer) reference modules."""
    here = os.path.dirname(os.path.abspath(__file__))
    oracle = os.path.join(os.path.dirname(here), "oracle")
    if oracle not in sys.path:
        sys.path.insert(0, oracle)
    import gtn  # noqa: F401  (the oracle shim)

    mods = {}
    for name in ("ctc", "asg", "stc", "transducer"):
        full = "_reference_criterions_" + name
        if full in sys.modules:
            mods[name] = sys.modules[full]
            continue
        spec = importlib.util.spec_from_file_location(
            full, os.path.join(REFERENCE, "criterions", name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[full] = mod
        spec.loader.exec_module(mod)
        mods[name] = mod
    return mods["ctc"], mods["asg"], mods["stc"], mods["transducer"]
