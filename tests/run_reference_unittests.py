"""Runs the reference's OWN unittest files, unchanged, against the oracle `gtn`
shim (build container only; /root/reference is read at run time).  Compat names
for the bit-rotted imports (`import transducer`, `from utils import CTCLoss ...`)
are injected as SURVEY.md Appendix C.1 describes.  Exit code 0 iff every
non-skipped reference test passes."""
import importlib.util
import os
import sys
import unittest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
REFERENCE = "/root/reference"
sys.path.insert(0, REFERENCE)

import gtn  # noqa: E402,F401
from criterions import asg, ctc, transducer as tr  # noqa: E402

sys.modules["transducer"] = tr
spec = importlib.util.spec_from_file_location("utils", os.path.join(REFERENCE, "utils.py"))
utils = importlib.util.module_from_spec(spec)
sys.modules["utils"] = utils
spec.loader.exec_module(utils)
utils.CTCLoss = ctc.CTCLoss
utils.ASGLoss = asg.ASGLoss
utils.ASGLossFunction = asg.ASGLossFunction
utils.pack_replabels = asg.pack_replabels
utils.unpack_replabels = asg.unpack_replabels

os.chdir(os.path.join(REFERENCE, "tests"))
names = sys.argv[1:] or ["gtn_ctc_test", "gtn_asg_test", "gtn_stc_test", "utils_test", "transducer_test"]
failed = 0
ran = 0
for name in names:
    spec = importlib.util.spec_from_file_location(name, os.path.join(REFERENCE, "tests", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    res = unittest.TextTestRunner(verbosity=1).run(unittest.defaultTestLoader.loadTestsFromModule(mod))
    ran += res.testsRun
    failed += len(res.failures) + len(res.errors)
print("REFERENCE_UNITTESTS ran=%d failed=%d" % (ran, failed))
sys.exit(1 if failed or ran == 0 else 0)
