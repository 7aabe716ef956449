"""The chain-split scaled CTC kernels through the C ABI — csrc/ctc_chain.cu (both directions
packed in f32x2; wfst_debug_force_generic_ctc(5)) and csrc/ctc_solo.cu (one direction per warp
set, scalar; hook 6) — forced ahead of the paired kernel, against the float64 DP: every (K, W)
configuration incl. more warps than the target needs (the warp-to-warp ring, the carry-in of the
exponent scan, partial steps next to the meeting point), ragged / empty targets, T down to 1, the
BASELINE shapes (cfg2 slice, cfg5 slice: the shape the paired layout cannot hold)."""
import numpy as np
import pytest
import torch

from _capi import ctc_capi

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[5, 6, 8], ids=["packed", "solo", "tick"])
def chain_first(request):
    from gtn_applications_b200 import _lib
    L = _lib.lib()
    old = L.wfst_debug_force_generic_ctc(request.param)
    yield L
    L.wfst_debug_ctc_chain_config(0, 0)
    L.wfst_debug_force_generic_ctc(old)


def check(B, T, C, L, seed=0, ragged=False, scale=1.0, nchk=4, max_fallback=0):
    import dp_numpy
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, T, C, generator=g) * scale
    if ragged:
        tg = [torch.randint(C - 1, (int(n),), generator=g).tolist() for n in torch.randint(0, L + 1, (B,), generator=g)]
    else:
        tg = torch.randint(C - 1, (B, L), generator=g).tolist()
    e = torch.log_softmax(x, 2)
    losses, mean, grad, flags = ctc_capi(e.cuda(), tg, C - 1)
    assert int((flags != 0).sum()) <= max_fallback, flags
    worst = 0.0
    for b in range(min(nchk, B)):
        Z, gZ = dp_numpy.ctc_dense_one(e[b].numpy().astype(np.float64), tg[b], C - 1)
        if not np.isfinite(Z):
            assert not np.isfinite(losses[b])
            continue
        assert abs(losses[b] + Z) <= 1e-4 * abs(Z) + 1e-6
        want = -gZ / B
        sc = np.abs(want).max()
        worst = max(worst, float((np.abs(grad[b] - want) / (1e-4 * np.abs(want) + 1e-4 * sc)).max()))
    assert worst <= 1.0, worst
    rows = grad.sum(2)
    fin = np.isfinite(losses)
    np.testing.assert_allclose(rows[fin], np.full_like(rows[fin], -1.0 / B), rtol=2e-4)


@pytest.mark.parametrize("cfg", [(0, 0), (4, 2), (4, 3), (6, 2), (6, 3), (4, 4), (6, 4)])
@pytest.mark.parametrize("shape", [(4, 50, 12, 7), (4, 1, 5, 0), (3, 17, 9, 8), (5, 100, 30, 40), (3, 16, 6, 3), (2, 33, 40, 16)])
def test_small_shapes_every_configuration(chain_first, cfg, shape):
    chain_first.wfst_debug_ctc_chain_config(*cfg)
    check(*shape)
    check(*shape, ragged=True, seed=1)


@pytest.mark.parametrize("shape", [(8, 300, 30, 100), (4, 333, 40, 150), (16, 1000, 30, 176), (8, 1500, 80, 264)])
def test_long_targets(chain_first, shape):
    check(*shape)


@pytest.mark.parametrize("shape", [(4, 777, 100, 300), (4, 640, 120, 383)])
def test_longest_targets_tight_in_time(chain_first, shape):
    """T barely above 2L: alignments are squeezed, some utterances exceed float32's range inside a
    lane and are handed to the float64 kernel; whoever computes them, the result holds"""
    check(*shape, max_fallback=shape[0])


def test_cfg2_full_batch_rows_and_slice(chain_first):
    check(256, 1000, 30, 176, seed=3, nchk=8)


def test_steep_emissions_are_flagged_and_recomputed_in_float64(chain_first):
    check(16, 1000, 30, 176, scale=3.0, max_fallback=16)


def test_target_with_blank_label_goes_to_the_fallback(chain_first):
    B, T, C = 2, 40, 8
    g = torch.Generator().manual_seed(5)
    e = torch.log_softmax(torch.randn(B, T, C, generator=g), 2)
    tg = [[1, C - 1, 2], [3, 4]]
    import dp_numpy
    losses, mean, grad, flags = ctc_capi(e.cuda(), tg, C - 1)
    assert flags[0] != 0 and flags[1] == 0
    Z, gZ = dp_numpy.ctc_dense_one(e[1].numpy().astype(np.float64), tg[1], C - 1)
    assert abs(losses[1] + Z) <= 1e-5 * abs(Z)
    np.testing.assert_allclose(grad[1], -gZ / B, atol=1e-6)


@pytest.mark.parametrize("shape", [(256, 1000, 30, 176), (5, 100, 30, 40), (8, 1500, 80, 264)])
def test_results_are_bit_identical_from_run_to_run(chain_first, shape):
    """Every hand-off between the warps of a block goes through an mbarrier; a missed one would
    show up as run-to-run differences (compute-sanitizer's racecheck cannot follow inline-PTX
    mbarriers, profiles/r2_sanitizer.txt).  No float atomics anywhere in these kernels: three
    runs of the same batch are bit-identical."""
    B, T, C, L = shape
    g = torch.Generator().manual_seed(11)
    e = torch.log_softmax(torch.randn(B, T, C, generator=g), 2).cuda()
    tg = torch.randint(C - 1, (B, L), generator=g).tolist()
    runs = [ctc_capi(e, tg, C - 1) for _ in range(3)]
    for r in runs[1:]:
        assert np.array_equal(r[0], runs[0][0]) and np.array_equal(r[2], runs[0][2]) and np.array_equal(r[3], runs[0][3])
