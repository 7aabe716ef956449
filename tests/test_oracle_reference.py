"""Pins the oracle to the reference (build container only — needs
/root/reference): (1) the reference's own unittest files pass UNCHANGED on the
oracle `gtn` shim; (2) oracle/ref_criterions.py (the restatement that travels to
the GPU box) is identical up to float32 rounding of the final scale/mean (<=2e-7 rel) to the reference's criterions/*.py run
on the same shim, on seeded random inputs."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.reference
HERE = os.path.dirname(os.path.abspath(__file__))


def test_reference_unittests_pass_on_shim():
    out = subprocess.run([sys.executable, os.path.join(HERE, "run_reference_unittests.py")],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "failed=0" in out.stdout


def _ragged(rng, B, C, lo, hi):
    return [rng.integers(0, C, size=rng.integers(lo, hi + 1)).tolist() for _ in range(B)]


@pytest.mark.parametrize("reduction", ["none", "mean"])
def test_ctc_restatement_equals_reference(gtn32, reduction):
    from _refload import load_reference
    import ref_criterions as rc
    ctc, _, _, _ = load_reference()
    rng = np.random.default_rng(0)
    B, T, C = 5, 17, 7
    x = torch.randn(B, T, C, generator=torch.Generator().manual_seed(1))
    lp = torch.log_softmax(x, 2).detach().requires_grad_(True)
    tg = _ragged(rng, B, C - 1, 0, 6)
    loss = ctc.CTCLoss(lp, tg, C - 1, reduction)
    loss.backward()
    mine = rc.ctc(gtn32, lp.detach().numpy(), tg, C - 1, reduction)
    assert abs(mine["loss"] - loss.item()) <= 2e-7 * abs(loss.item())  # fp32 mean vs fp64 mean
    np.testing.assert_allclose(mine["grad"], lp.grad.numpy(), rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("reduction", ["none", "mean"])
def test_asg_restatement_equals_reference(gtn32, reduction):
    from _refload import load_reference
    import ref_criterions as rc
    _, asg, _, _ = load_reference()
    rng = np.random.default_rng(2)
    B, T, C = 4, 13, 6
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, T, C, generator=g, requires_grad=True)
    tr = torch.randn(C + 1, C, generator=g, requires_grad=True)
    tg = _ragged(rng, B, C, 1, 6)
    loss = asg.ASGLoss(x, tr, tg, reduction)
    loss.backward()
    mine = rc.asg(gtn32, x.detach().numpy(), tr.detach().numpy(), tg, reduction)
    assert abs(mine["loss"] - loss.item()) <= 2e-7 * abs(loss.item())  # fp32 mean vs fp64 mean
    np.testing.assert_allclose(mine["grad"], x.grad.numpy(), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(mine["grad_transitions"], tr.grad.numpy(), rtol=1e-6, atol=1e-7)


def test_stc_restatement_equals_reference(gtn32):
    from _refload import load_reference
    import ref_criterions as rc
    _, _, stc, _ = load_reference()
    rng = np.random.default_rng(4)
    B, T, C = 3, 11, 5
    x = torch.randn(B, T, 2 * C, generator=torch.Generator().manual_seed(5), requires_grad=True)
    tg = [rng.integers(1, C, size=rng.integers(0, 4)).tolist() for _ in range(B)]
    for reduction in ("none", "mean"):
        x.grad = None
        loss = stc.STCLoss(x, tg, 0.7, reduction)
        loss.backward()
        mine = rc.stc(gtn32, x.detach().numpy(), tg, 0.7, reduction)
        assert abs(mine["loss"] - loss.item()) <= 2e-7 * abs(loss.item())  # fp32 mean vs fp64 mean
        np.testing.assert_allclose(mine["grad"], x.grad.numpy(), rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("blank,allow_repeats", [("none", True), ("optional", True),
                                                 ("optional", False), ("forced", True)])
def test_transducer_restatement_equals_reference(gtn32, blank, allow_repeats):
    from _refload import load_reference
    import ref_criterions as rc
    _, _, _, tr = load_reference()
    tokens = ["a", "b", "ab", "ba", "aba"]
    g2i = {"a": 0, "b": 1}
    C = len(tokens) + int(blank != "none")
    B, T = 3, 9
    x = torch.randn(B, T, C, generator=torch.Generator().manual_seed(6), requires_grad=True)
    tg = [[0, 1, 0], [1, 1, 0, 1], [0]]
    ref = tr.Transducer(tokens, g2i, blank=blank, allow_repeats=allow_repeats, reduction="mean")
    loss = ref(x, tg)
    loss.backward()
    mine = rc.Transducer(gtn32, tokens, g2i, blank=blank, allow_repeats=allow_repeats, reduction="mean")
    lsm = torch.log_softmax(x.detach(), 2).numpy()
    res = mine.loss(lsm, tg)
    assert abs(res["loss"] - loss.item()) < 1e-6
    # reference gradient is w.r.t. the logits; push ours through log_softmax
    g = torch.from_numpy(res["grad"]).float()
    x2 = x.detach().clone().requires_grad_(True)
    torch.log_softmax(x2, 2).backward(g)
    np.testing.assert_allclose(x2.grad.numpy(), x.grad.numpy(), rtol=1e-5, atol=1e-6)
    # structures: token / lexicon graphs bit-exact
    for a, b in ((ref.tokens, mine.tokens), (ref.lexicon, mine.lexicon)):
        ga, gb = rc.graph_arrays(a), rc.graph_arrays(b)
        for k in ga:
            np.testing.assert_array_equal(ga[k], gb[k])
    assert mine.viterbi(lsm) == [p.tolist() for p in ref.viterbi(torch.from_numpy(lsm))]


def test_transducer_backoff_transitions_equals_reference(gtn32):
    from _refload import load_reference
    import ref_criterions as rc
    _, _, _, tr = load_reference()
    N, T = 5, 6
    tokens = [(n,) for n in range(N)]
    g2i = {n: n for n in range(N)}
    path = "/root/reference/tests/trans_backoff_test.txt"
    x = torch.randn(2, T, N + 1, generator=torch.Generator().manual_seed(7), requires_grad=True)
    tg = [[0, 1, 0], [2, 2]]
    ref = tr.Transducer(tokens, g2i, blank="optional", allow_repeats=False,
                        transitions=gtn32.loadtxt(path))
    ref.transition_params.data = torch.randn(ref.transition_params.numel(),
                                             generator=torch.Generator().manual_seed(8)) * 0.3
    loss = ref(x, tg)
    loss.backward()
    mine = rc.Transducer(gtn32, tokens, g2i, blank="optional", allow_repeats=False,
                         transitions=gtn32.loadtxt(path))
    res = mine.loss(x.detach().numpy(), tg, ref.transition_params.detach().numpy())
    assert abs(res["loss"] - loss.item()) < 1e-6
    np.testing.assert_allclose(res["grad"], x.grad.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(res["grad_transitions"], ref.transition_params.grad.numpy(),
                               rtol=1e-5, atol=1e-6)
