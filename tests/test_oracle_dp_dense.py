"""The vectorised float64 CTC recursion (oracle/dp_numpy.py: ctc_dense) that the BASELINE-size
GPU parity tests use as truth equals the arc-list DP (dp_numpy.ctc) and the oracle's GTN
restatement on small cases, including empty targets, repeats and infeasible alignments."""
import math

import numpy as np
import pytest

import dp_numpy
import _golden as G


@pytest.mark.parametrize("B,T,C,lens,lsm", [
    (3, 12, 6, [4, 0, 6], True), (2, 64, 5, [30, 32], False), (4, 40, 7, [1, 2, 19, 20], True),
    (2, 5, 4, [3, 1], False), (1, 1, 3, [0], True), (2, 3, 4, [5, 1], True)])
def test_dense_recursion_equals_arc_list_dp(B, T, C, lens, lsm):
    rng = np.random.default_rng(B * 100 + T)
    E = rng.standard_normal((B, T, C))
    if lsm:
        E = G.log_softmax(E)
    tg = [rng.integers(0, C - 1, size=n).tolist() for n in lens]
    tg[0] = [tg[0][0]] * len(tg[0]) if tg[0] else tg[0]        # a run of repeats
    a = dp_numpy.ctc(E, tg, C - 1, "mean")
    b = dp_numpy.ctc_dense(E, tg, C - 1, "mean")
    for x, y in zip(a["losses"], b["losses"]):
        assert (math.isinf(x) and math.isinf(y)) or abs(x - y) <= 1e-12 * max(1.0, abs(x))
    np.testing.assert_allclose(b["grad"], a["grad"], rtol=1e-10, atol=1e-14)


def test_dense_recursion_equals_gtn_restatement(gtn64):
    import ref_criterions as rc
    rng = np.random.default_rng(5)
    E = G.log_softmax(rng.standard_normal((3, 30, 8))).astype(np.float32)
    tg = [[1, 1, 2], [], [0, 3, 3, 5, 6, 6, 2]]
    ref = rc.ctc(gtn64, E, tg, 7, "mean")
    got = dp_numpy.ctc_dense(E, tg, 7, "mean")
    assert abs(got["loss"] - ref["loss"]) <= 1e-9 * abs(ref["loss"])
    np.testing.assert_allclose(got["grad"], ref["grad"], rtol=1e-8, atol=1e-12)
