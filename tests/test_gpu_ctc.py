"""GPU parity: the CUDA CTC path (through CTCLossFunction -> C ABI) against the
float64 oracle on the same seeded inputs, the reference's literal vectors, the
committed fixtures, and size-independent properties at BASELINE.json's sizes.
Tolerance (north_star: "within 1e-4 relative (fp32)"): loss within 1e-4 relative;
gradient |a-b| <= 1e-4 * |b| + 1e-4 * max|b| elementwise, i.e. rtol 1e-4 plus an
absolute floor of 1e-4 of the tensor's max-norm — the shape of the reference's own
comparison (rtol=1e-4, atol=1e-5 on O(0.1) gradients, tests/transducer_test.py:315).
The truth is the float64 oracle: float32 GTN arithmetic itself is only good to ~5e-3
at T=1000 ("Numerics")."""
import math

import numpy as np
import pytest
import torch

import _golden as G

pytestmark = pytest.mark.gpu


def assert_close(got, want, rel=1e-4):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    scale = np.abs(want).max() if want.size else 0.0
    tol = rel * np.abs(want) + rel * scale + 1e-12
    bad = np.abs(got - want) > tol
    assert not bad.any(), "max abs err %.3e (scale %.3e), %d bad" % (
        np.abs(got - want).max(), scale, int(bad.sum()))


def assert_close_f32_fixture(got, want):
    """Fixtures hold the reference's float32 GTN arithmetic (own error ~1e-6 abs on
    O(0.1) posteriors): rtol 1e-3 with an absolute floor of 1e-5 * max|want|."""
    want = np.asarray(want, dtype=np.float64)
    np.testing.assert_allclose(np.asarray(got, dtype=np.float64), want, rtol=1e-3,
                               atol=1e-5 * max(np.abs(want).max(), 1e-30))


def run(lp_np, targets, blank, reduction):
    from gtn_applications_b200.criterions.ctc import CTCLoss
    lp = torch.tensor(lp_np, dtype=torch.float32, device="cuda").requires_grad_(True)
    loss = CTCLoss(lp, targets, blank, reduction)
    loss.backward()
    return loss.item(), lp.grad.cpu().numpy()


def test_trivial_and_uniform():
    with np.errstate(divide="ignore"):
        lp = np.log(np.array([1.0, 0.0, 0.0, 1.0, 1.0, 0.0]).reshape(1, 3, 2))
    loss, _ = run(lp, [[0, 0]], 1, "none")
    assert abs(loss) < 1e-6
    loss, _ = run(G.log_softmax(np.zeros((1, 3, 4))), [[1, 2]], 3, "none")
    assert abs(loss + math.log(0.25 ** 3 * 5)) < 1e-5


def test_warpctc_vectors():
    import test_oracle_golden as lit
    for probs, labels, want, want_grad in ((lit.WARP_CTC_1, [[0, 1, 2, 1, 0]], 3.34211, lit.WARP_CTC_1_GRAD),
                                           (lit.WARP_CTC_2, [[0, 1, 1, 0]], 5.42262, lit.WARP_CTC_2_GRAD)):
        logits = np.log(probs)
        loss, grad = run(G.log_softmax(logits), labels, 5, "none")
        assert abs(loss - want) < 5e-5
        np.testing.assert_allclose(G.through_log_softmax(logits, grad), want_grad, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("case", ["small_none", "small_mean", "raw_mean", "repeats", "cfg1_raw", "cfg1_lsm"])
def test_fixtures_from_reference(case):
    z = G.load("ctc")
    tg = G.unpack(z[case + "_targets"], z[case + "_offsets"])
    loss, grad = run(z[case + "_emissions"], tg, int(z[case + "_blank"]), str(z[case + "_reduction"]))
    want = float(z[case + "_loss"])
    assert abs(loss - want) <= 1e-4 * max(1.0, abs(want))
    # the fixture itself is float32 GTN arithmetic: compare at 1e-3 here, 1e-4 vs float64 below
    assert_close_f32_fixture(grad, z[case + "_grad"])


@pytest.mark.parametrize("B,T,C,lens,lsm,reduction", [
    (3, 12, 6, [4, 0, 6], True, "none"),
    (5, 33, 9, [1, 16, 7, 0, 11], True, "mean"),
    (4, 150, 28, [20, 20, 20, 20], False, "none"),     # BASELINE configs[0]
    (2, 64, 5, [30, 32], False, "mean"),                # T barely enough (repeats may make it infeasible)
    (6, 257, 31, [40, 3, 77, 128, 0, 64], True, "mean"),
])
def test_against_float64_oracle(gtn64, B, T, C, lens, lsm, reduction):
    import ref_criterions as rc
    rng = np.random.default_rng(B * 1000 + T)
    x = rng.standard_normal((B, T, C)).astype(np.float32)
    lp = G.log_softmax(x).astype(np.float32) if lsm else x
    tg = [rng.integers(0, C - 1, size=n).tolist() for n in lens]
    ref = rc.ctc(gtn64, lp, tg, C - 1, reduction)
    loss, grad = run(lp, tg, C - 1, reduction)
    if math.isinf(ref["loss"]):
        assert math.isinf(loss)
        return
    assert abs(loss - ref["loss"]) <= 1e-4 * abs(ref["loss"])
    assert_close(grad, ref["grad"])


def test_infeasible_alignment_gives_inf_loss_and_zero_grad():
    lp = G.log_softmax(np.random.default_rng(0).standard_normal((2, 3, 4))).astype(np.float32)
    loss, grad = run(lp, [[0, 1, 2, 0, 1], [1]], 3, "none")
    assert math.isinf(loss) and loss > 0
    assert np.all(grad[0] == 0) and np.isfinite(grad).all() and np.abs(grad[1]).sum() > 0


def test_errors_match_reference():
    from gtn_applications_b200.criterions.ctc import CTCLoss
    lp = torch.zeros(1, 4, 3, device="cuda")
    with pytest.raises(ValueError, match="invalid value for reduction"):
        CTCLoss(lp, [[0]], 2, "sum")
    with pytest.raises(TypeError):
        CTCLoss(lp.double(), [[0]], 2, "none")
    with pytest.raises(ValueError):
        CTCLoss(lp, [[7]], 2, "none")


def test_grad_output_scaling_and_cpu_input():
    from gtn_applications_b200.criterions.ctc import CTCLoss
    rng = np.random.default_rng(3)
    lp0 = torch.tensor(G.log_softmax(rng.standard_normal((2, 20, 5))), dtype=torch.float32)
    tg = [[0, 1, 2], [3, 3]]
    a = lp0.clone().cuda().requires_grad_(True)
    CTCLoss(a, tg, 4, "mean").backward()
    b = lp0.clone().cuda().requires_grad_(True)
    (CTCLoss(b, tg, 4, "mean") * 2.5).backward()
    torch.testing.assert_close(b.grad, a.grad * 2.5)
    c = lp0.clone().requires_grad_(True)  # host tensor: result and grad come back on the host
    out = CTCLoss(c, tg, 4, "mean")
    out.backward()
    assert not out.is_cuda and not c.grad.is_cuda
    torch.testing.assert_close(c.grad, a.grad.cpu())


@pytest.mark.parametrize("dtype", [torch.int64, torch.int32])
@pytest.mark.parametrize("reduction", ["none", "mean"])
def test_device_resident_rectangular_targets(dtype, reduction):
    """A [B, L] label tensor that already lives on the device is used in place (no host round
    trip): same loss and gradient as the list-of-lists call."""
    from gtn_applications_b200.criterions.ctc import CTCLoss
    torch.manual_seed(3)
    B, T, C, L = 6, 90, 13, 17
    lp0 = torch.log_softmax(torch.randn(B, T, C, device="cuda"), 2)
    tg = torch.randint(C - 1, (B, L))
    a = lp0.clone().requires_grad_(True)
    la = CTCLoss(a, tg.tolist(), C - 1, reduction)
    la.backward()
    for _ in range(2):                      # second call: cached offsets / scales
        b = lp0.clone().requires_grad_(True)
        lb = CTCLoss(b, tg.to("cuda", dtype), C - 1, reduction)
        lb.backward()
        assert lb.item() == la.item()
        assert torch.equal(a.grad, b.grad)


def test_full_size_properties():
    """BASELINE configs[1] (B=256, T=1000, C=30, L=176): size-independent checks —
    every frame's posteriors sum to one (so grad rows sum to -scale/B), the
    gradient is non-positive, and loss equals torch's own ctc_loss."""
    from gtn_applications_b200.criterions.ctc import CTCLoss
    torch.manual_seed(0)
    B, T, C, L = 256, 1000, 30, 176
    lp = torch.log_softmax(torch.randn(B, T, C, device="cuda"), 2).requires_grad_(True)
    tg = torch.randint(C - 2, (B, L))
    loss = CTCLoss(lp, tg.tolist(), C - 1, "none")
    loss.backward()
    rows = lp.grad.sum(2)
    assert torch.all(lp.grad <= 1e-7)
    torch.testing.assert_close(rows, torch.full_like(rows, -1.0 / B), rtol=2e-4, atol=0)
    ref = torch.nn.functional.ctc_loss(lp.detach().permute(1, 0, 2), tg.cuda(), [T] * B, [L] * B,
                                       blank=C - 1, reduction="none").mean()
    assert abs(loss.item() - ref.item()) <= 1e-4 * abs(ref.item())


# --------------------------------------------------------------------------- kernels / dispatch
def _hazards(B, T, C, L):
    from gtn_applications_b200 import _lib, _runtime as rt
    flags = np.zeros(B, dtype=np.int32)
    ws = rt.workspace(torch.device("cuda:0"), _lib.lib().wfst_ctc_workspace_bytes(B, T, C, L))
    _lib.check(_lib.lib().wfst_debug_ctc_hazards(ws.data_ptr(), B, T, C, L, flags.ctypes.data))
    return flags


@pytest.mark.parametrize("kind", [0, 2, 1])   # default (paired kernel), single-utterance scaled kernel, log-semiring only
@pytest.mark.parametrize("B,T,C,lens,scale", [
    (5, 203, 30, [40, 3, 77, 95, 0], 1.0),          # odd B: the last pair block runs one utterance twice
    (4, 160, 12, [9, 30, 1, 22], 1.0),
    (3, 120, 30, [30, 30, 30], 6.0),                # steep scores: utterances leave the scaled kernels
    (2, 96, 70, [20, 47], 1.0),                     # C > 31: wider p tiles
])
def test_every_ctc_kernel_against_numpy_dp(kind, B, T, C, lens, scale, lattice_kernel):
    """The three CTC kernels (and the hand-off between them) against the closed-form float64 DP."""
    import dp_numpy
    from gtn_applications_b200 import _lib
    rng = np.random.default_rng(7 * B + T)
    lp = G.log_softmax(rng.standard_normal((B, T, C)) * scale).astype(np.float32)
    tg = [rng.integers(0, C - 1, size=n).tolist() for n in lens]
    ref = dp_numpy.ctc(lp, tg, C - 1, "mean")
    old = _lib.lib().wfst_debug_force_generic_ctc(kind)
    try:
        loss, grad = run(lp, tg, C - 1, "mean")
    finally:
        _lib.lib().wfst_debug_force_generic_ctc(old)
    assert abs(loss - ref["loss"]) <= 1e-4 * abs(ref["loss"])
    assert_close(grad, ref["grad"])


def test_paired_kernel_is_the_default_and_keeps_the_benchmark_shape():
    """At the benchmark's distribution no utterance may fall back to the log-semiring kernel."""
    from gtn_applications_b200.criterions.ctc import CTCLoss
    torch.manual_seed(1)
    B, T, C, L = 16, 1000, 30, 176
    lp = torch.log_softmax(torch.randn(B, T, C, device="cuda"), 2).requires_grad_(True)
    CTCLoss(lp, torch.randint(C - 2, (B, L)).tolist(), C - 1, "none").backward()
    torch.cuda.synchronize()
    assert (_hazards(B, T, C, L) == 0).all()


@pytest.mark.parametrize("B,T,C,lens,scale", [
    (4, 150, 28, [20, 20, 20, 20], 1.0),
    (5, 203, 32, [40, 3, 77, 95, 0], 1.0),          # odd B, partial last tile
    (3, 120, 30, [30, 30, 30], 6.0),                # flagged utterances: log-softmax fallback chain
    (2, 64, 8, [40, 3], 1.0),                       # an infeasible alignment rides with a feasible one
])
def test_fused_log_softmax_matches_two_step_path_and_oracle(B, T, C, lens, scale):
    """CTCLogitsLoss(x) == CTCLoss(log_softmax(x)) in value and in d/dx (criterions/ctc.py:107,
    tests/gtn_ctc_test.py:64-80), and both match the float64 DP pushed through the softmax."""
    import dp_numpy
    from gtn_applications_b200.criterions.ctc import CTCLoss, CTCLogitsLoss, CTCLogitsLossFunction
    rng = np.random.default_rng(11 * B + T)
    x = (rng.standard_normal((B, T, C)) * scale).astype(np.float32)
    tg = [rng.integers(0, C - 1, size=n).tolist() for n in lens]
    a = torch.tensor(x, device="cuda").requires_grad_(True)
    assert CTCLogitsLossFunction.supported(a, tg)
    la = CTCLogitsLoss(a, tg, C - 1, "mean")
    la.backward()
    b = torch.tensor(x, device="cuda").requires_grad_(True)
    lb = CTCLoss(torch.log_softmax(b, 2), tg, C - 1, "mean")
    lb.backward()
    ref = dp_numpy.ctc(G.log_softmax(x), tg, C - 1, "mean")
    if math.isinf(ref["loss"]):
        assert math.isinf(la.item()) and math.isinf(lb.item())
    else:
        assert abs(la.item() - ref["loss"]) <= 1e-4 * abs(ref["loss"])
        assert abs(la.item() - lb.item()) <= 1e-5 * abs(lb.item())
    want = G.through_log_softmax(x, ref["grad"])
    assert_close(a.grad.cpu().numpy(), want)
    assert_close(b.grad.cpu().numpy(), want)


def test_ctc_module_uses_fused_path_and_matches_torch():
    from gtn_applications_b200.criterions.ctc import CTC
    torch.manual_seed(3)
    B, T, C = 6, 200, 30
    x = torch.randn(B, T, C, device="cuda")
    targets = [torch.randint(C - 1, (n,)) for n in (30, 12, 44, 1, 25, 60)]
    a = x.clone().requires_grad_(True)
    la = CTC(C - 1, False)(a, targets)
    la.backward()
    b = x.clone().requires_grad_(True)
    lb = CTC(C - 1, True)(b, targets)          # torch's ctc_loss: mean over (loss_b / L_b), like reduction="mean"
    lb.backward()
    assert abs(la.item() - lb.item()) <= 1e-4 * abs(lb.item())
    torch.testing.assert_close(a.grad, b.grad, rtol=1e-3, atol=1e-6)
