"""GPU parity: the CUDA ASG path (ASGLossFunction -> wfst_asg_forward_backward)
against the float64 oracle, the reference's known-answer vector and the committed
fixtures.  Tolerance 1e-4 relative (north_star)."""
import math

import numpy as np
import pytest
import torch

import _golden as G
from test_gpu_ctc import assert_close, assert_close_f32_fixture

pytestmark = pytest.mark.gpu


def run(e_np, tr_np, targets, reduction):
    from gtn_applications_b200.criterions.asg import ASGLoss
    e = torch.tensor(e_np, dtype=torch.float32, device="cuda").requires_grad_(True)
    tr = torch.tensor(tr_np, dtype=torch.float32, device="cuda").requires_grad_(True)
    loss = ASGLoss(e, tr, targets, reduction)
    loss.backward()
    return loss.item(), e.grad.cpu().numpy(), tr.grad.cpu().numpy()


def test_known_answer():
    import test_oracle_golden as lit
    loss, ge, gt = run(lit.ASG_EMISSIONS, np.zeros((7, 6)), lit.ASG_LABELS, "none")
    assert abs(loss - 7.47995) < 5e-5
    np.testing.assert_allclose(ge, lit.ASG_GRAD, rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(gt[1:], lit.ASG_TRANS_GRAD, rtol=1e-3, atol=1e-5)


@pytest.mark.parametrize("case", ["small_none", "small_mean", "mid_mean"])
def test_fixtures_from_reference(case, lattice_kernel):
    z = G.load("asg")
    tg = G.unpack(z[case + "_targets"], z[case + "_offsets"])
    loss, ge, gt = run(z[case + "_emissions"], z[case + "_transitions"], tg, str(z[case + "_reduction"]))
    want = float(z[case + "_loss"])
    assert abs(loss - want) <= 1e-4 * max(1.0, abs(want))
    assert_close_f32_fixture(ge, z[case + "_grad"])
    assert_close_f32_fixture(gt, z[case + "_grad_transitions"])


@pytest.mark.parametrize("B,T,C,lens,reduction", [
    (3, 10, 5, [3, 5, 1], "none"),
    (4, 40, 12, [9, 14, 2, 20], "mean"),
    (2, 120, 30, [40, 60], "mean"),
    (3, 64, 80, [10, 31, 5], "none"),     # reference benchmark's N=80 (asg_benchmark.py:19)
])
def test_against_float64_oracle(gtn64, B, T, C, lens, reduction, lattice_kernel):
    import ref_criterions as rc
    rng = np.random.default_rng(B * 100 + T)
    e = rng.standard_normal((B, T, C)).astype(np.float32)
    tr = rng.standard_normal((C + 1, C)).astype(np.float32)
    tg = [rng.integers(0, C, size=n).tolist() for n in lens]
    ref = rc.asg(gtn64, e, tr, tg, reduction)
    loss, ge, gt = run(e, tr, tg, reduction)
    assert abs(loss - ref["loss"]) <= 1e-4 * abs(ref["loss"])
    assert_close(ge, ref["grad"])
    assert_close(gt, ref["grad_transitions"])


def test_module_replabels_garbage_and_checkpoint_names():
    from gtn_applications_b200.criterions.asg import ASG
    z = G.load("asg")
    crit = ASG(4, num_replabels=2, use_garbage=True).cuda()
    assert list(crit.state_dict().keys()) == ["transitions"]
    assert tuple(crit.transitions.shape) == (crit.N + 1, crit.N)
    crit.transitions.data = torch.tensor(z["module_transitions"], device="cuda")
    x = torch.tensor(z["module_emissions"], device="cuda", requires_grad=True)
    tg = [torch.tensor(t) for t in G.unpack(z["module_targets"], z["module_offsets"])]
    loss = crit(x, tg)
    loss.backward()
    assert abs(loss.item() - float(z["module_loss"])) <= 1e-4 * abs(float(z["module_loss"]))
    assert_close_f32_fixture(x.grad.cpu().numpy(), z["module_grad"])
    assert_close_f32_fixture(crit.transitions.grad.cpu().numpy(), z["module_grad_transitions"])


@pytest.mark.parametrize("B,T,C,lens", [(5, 77, 30, [20, 1, 33, 8, 40]), (2, 9, 3, [2, 4]), (3, 130, 32, [5, 60, 17]),
                                        (1, 8, 30, [3]), (3, 5, 7, [1, 2, 5])])
def test_dense_and_generic_full_connect_kernels_agree(B, T, C, lens):
    """The dense warp-per-utterance full-connect kernel (C <= 32) and the generic lattice
    kernel compute the same loss and gradients (the generic path is forced with the CTC hook)."""
    from gtn_applications_b200 import _lib
    rng = np.random.default_rng(B + T + C)
    e = (rng.standard_normal((B, T, C)) * 2).astype(np.float32)
    tr = rng.standard_normal((C + 1, C)).astype(np.float32)
    tg = [rng.integers(0, C, size=n).tolist() for n in lens]
    dense = run(e, tr, tg, "mean")           # two warps per utterance that meet in the middle (T >= 8)
    old = _lib.lib().wfst_debug_force_generic_ctc(3)
    try:
        single = run(e, tr, tg, "mean")      # one warp per utterance
        _lib.lib().wfst_debug_force_generic_ctc(1)
        generic = run(e, tr, tg, "mean")
    finally:
        _lib.lib().wfst_debug_force_generic_ctc(old)
    for got in (dense, single):
        assert abs(got[0] - generic[0]) <= 1e-5 * abs(generic[0])
        assert_close(got[1], generic[1])
        assert_close(got[2], generic[2])


def test_full_size_properties_cfg3():
    """BASELINE configs[2] (ASG B=256, T=1000, C=30, L=176): size-independent checks.  Both
    lattices cross every frame exactly once, so the emission-gradient rows sum to zero
    (full-connect posteriors minus force-align posteriors) and so does the transition gradient
    (T arcs taken in each lattice); the force-align term is the same on the shared-memory kernel
    and on the generic kernel; a slice of utterances agrees with the float64 DP."""
    import dp_numpy
    from gtn_applications_b200 import _lib
    from gtn_applications_b200.criterions.asg import ASGLoss
    torch.manual_seed(0)
    B, T, C, L = 256, 1000, 30, 176
    e0 = torch.randn(B, T, C, device="cuda")
    tr0 = torch.randn(C + 1, C, device="cuda")
    tg = torch.randint(C, (B, L)).tolist()

    def go():
        e, tr = e0.clone().requires_grad_(True), tr0.clone().requires_grad_(True)
        loss = ASGLoss(e, tr, tg, "none")
        loss.backward()
        return loss.item(), e.grad, tr.grad

    loss, ge, gt = go()
    assert math.isfinite(loss)
    assert float(ge.sum(2).abs().max()) <= 2e-4 / B
    assert abs(float(gt.sum())) <= 3e-4 * T      # +T and -T per utterance, accumulated in float32
    old = _lib.lib().wfst_debug_force_generic_lattice(1)
    try:
        loss2, ge2, gt2 = go()
    finally:
        _lib.lib().wfst_debug_force_generic_lattice(old)
    assert abs(loss - loss2) <= 1e-5 * abs(loss2)
    # the force-align term of the default path is the scaled-probability chain kernel (float64
    # slice below: within 1e-4); the generic log-semiring kernel keeps float32 log values of
    # magnitude ~3000 and is itself only good to a few 1e-4 of the max-norm at this T
    # (measured 2.4e-4): the cross check between the two carries both errors
    assert_close(ge.cpu().numpy(), ge2.cpu().numpy(), rel=5e-4)
    assert_close(gt.cpu().numpy(), gt2.cpu().numpy(), rel=5e-4)
    ref = dp_numpy.asg(e0[:3].cpu().numpy(), tr0.cpu().numpy(), tg[:3], "none")
    e = e0[:3].clone().requires_grad_(True)
    tr = tr0.clone().requires_grad_(True)
    l3 = ASGLoss(e, tr, tg[:3], "none")
    l3.backward()
    assert abs(l3.item() - ref["loss"]) <= 1e-4 * abs(ref["loss"])
    assert_close(e.grad.cpu().numpy(), ref["grad"])
    assert_close(tr.grad.cpu().numpy(), ref["grad_transitions"])


@pytest.mark.parametrize("B,T,C,lens,reduction", [
    (4, 2, 6, [1, 2, 1, 2], "none"),          # the shortest lattices the chain kernel takes (T >= 2)
    (3, 17, 9, [8, 17, 3], "none"),           # partial steps on both sides of the meeting point; T = L
    (4, 100, 10, [12, 1, 30, 7], "mean"),
    (2, 250, 80, [44, 44], "none"),           # benchmarks/asg_benchmark.py shapes
    (2, 333, 40, [190, 150], "mean"),         # 190 + 2 nodes: the last one-warp chain
    (2, 400, 30, [250, 320], "none"),         # two warps per chain (the ring between them)
    (2, 500, 30, [383, 5], "none"),           # the longest target the chain kernel holds next to a short one
    (2, 45, 12, [40, 50], "none"),            # one infeasible utterance (T < L): loss inf as in the other kernels
])
def test_force_align_chain_kernel_against_float64(B, T, C, lens, reduction):
    """csrc/asg_fal_chain.cu (scaled-probability force-align chain, the default for L <= 382):
    loss, emission gradient and transition gradient against the float64 DP, including the shapes
    that exercise its corners; the same inputs through the log-semiring lattice kernel."""
    import dp_numpy
    from gtn_applications_b200 import _lib
    rng = np.random.default_rng(7 * B + T)
    e = rng.standard_normal((B, T, C)).astype(np.float32)
    tr = rng.standard_normal((C + 1, C)).astype(np.float32)
    tg = [rng.integers(0, C, size=n).tolist() for n in lens]
    feasible = [b for b in range(B) if len(tg[b]) <= T]
    ref = dp_numpy.asg(e[feasible], tr, [tg[b] for b in feasible], "none")
    got = run(e[feasible], tr, [tg[b] for b in feasible], "none")
    assert abs(got[0] - ref["loss"]) <= 1e-4 * abs(ref["loss"])
    assert_close(got[1], ref["grad"])
    assert_close(got[2], ref["grad_transitions"])
    old = _lib.lib().wfst_debug_force_generic_lattice(2)     # log-semiring lean kernel for the force-align term
    try:
        lat = run(e, tr, tg, reduction)
    finally:
        _lib.lib().wfst_debug_force_generic_lattice(old)
    both = run(e, tr, tg, reduction)
    if len(feasible) == B:
        assert abs(both[0] - lat[0]) <= 1e-4 * abs(lat[0])
        assert_close(both[1], lat[1], rel=5e-4)      # carries the float32 log-semiring kernel's own error
        assert_close(both[2], lat[2], rel=5e-4)
    else:
        assert math.isinf(both[0]) and math.isinf(lat[0])
