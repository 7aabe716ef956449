"""The names the reference's callers take from `utils` (benchmarks/ctc_benchmark.py:13,
benchmarks/asg_benchmark.py:13, tests/transducer_test.py:19) on the product's criteria.
Models, datasets and the training loop of the reference's utils.py are outside the hot path
 and are not provided."""
from gtn_applications_b200.criterions.asg import ASG, ASGLoss, ASGLossFunction  # noqa: F401
from gtn_applications_b200.criterions.ctc import CTC, CTCLoss, CTCLossFunction  # noqa: F401
from gtn_applications_b200.criterions.stc import STC, STCLoss, STCLossFunction  # noqa: F401
from gtn_applications_b200.criterions.transducer import Transducer, TransducerLoss, TransducerLossFunction  # noqa: F401
