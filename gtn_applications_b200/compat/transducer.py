"""`import transducer` (reference benchmarks/transducer_benchmark.py:13, tests/transducer_test.py:17-18)."""
from gtn_applications_b200.criterions.transducer import *  # noqa: F401,F403
from gtn_applications_b200.criterions.transducer import (  # noqa: F401
    ConvTransduce1D, Transducer, TransducerLoss, TransducerLossFunction, make_chain_graph,
    make_kernel_graph, make_lexicon_graph, make_token_graph, make_transitions_graph)
