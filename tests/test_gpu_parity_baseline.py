"""GPU parity AT THE BASELINE SHAPES (BASELINE.json configs[1], [3], [4]):
the kernels that produce the bench numbers, called through the C ABI, compared ELEMENTWISE
with the float64 oracle on the same seeded inputs.

  * cfg2  CTC B=256, T=1000, C=30, L=176: every utterance's loss and the whole [B, T, C]
          gradient, for log_softmax(randn) emissions AND raw randn emissions (what
          benchmarks/ctc_benchmark.py:22,28 feeds), for the fused logits entry point, and for
          ragged targets L_b ~ U{100..250}; the compared utterances must not have been handed
          to the log-semiring fallback (hazard flag 0), so the comparison is of the scaled
          kernel itself.
  * cfg5  CTC T=1500, C=80, L=264 (per-GPU shard B=256): a slice of utterances.
  * cfg4  transducer on the reference's own word-piece list
          (tests/golden/word_pieces_tokens_1000.txt = benchmarks/word_pieces_tokens_1000.txt,
          transducer_benchmark.py:19-27,42-44): alignment graph bit-exact with the oracle's GTN
          restatement, loss and gradient of one utterance against the float64 DP.

Tolerance (north_star "within 1e-4 relative"): loss 1e-4 relative; gradient
|a-b| <= 1e-4*|b| + 1e-4*max|b| (see tests/test_gpu_ctc.py).  The measured maxima are appended
to gpurun_out/r2_parity.jsonl (copied to profiles/ by hand)."""
import json
import os
import random

import numpy as np
import pytest
import torch

import _golden as G
from _capi import ctc_capi

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def record(name, **vals):
    try:
        d = os.path.join(ROOT, "gpurun_out")
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "r2_parity.jsonl"), "a") as f:
            f.write(json.dumps(dict(case=name, **vals)) + "\n")
    except OSError:
        pass


def _dp_one(args):
    import dp_numpy
    E, y, blank = args
    return dp_numpy.ctc_dense_one(E, y, blank)


def dp_batch(E, targets, blank, idx):
    """float64 DP of the utterances `idx` on the host cores -> (logZ [n], dZ/dE [n, T, C])"""
    from concurrent.futures import ProcessPoolExecutor
    import multiprocessing as mp
    jobs = [(np.asarray(E[b], dtype=np.float64), list(targets[b]), blank) for b in idx]
    try:
        with ProcessPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1),
                                 mp_context=mp.get_context("fork")) as ex:
            res = list(ex.map(_dp_one, jobs, chunksize=4))
    except Exception:
        res = [_dp_one(j) for j in jobs]
    return np.array([r[0] for r in res]), np.stack([r[1] for r in res])


def compare(name, losses, grad, flags, E, targets, blank, idx, scales, softmax_of=None, max_fallback=0):
    """losses / grad / flags from the C ABI; oracle on utterances idx."""
    B = len(targets)
    Z, gZ = dp_batch(E, targets, blank, idx)
    want_loss = -Z
    got_loss = losses[idx]
    rel = np.abs(got_loss - want_loss) / np.abs(want_loss)
    want_grad = -gZ * np.asarray(scales, dtype=np.float64)[idx, None, None]
    if softmax_of is not None:   # d/d logits
        want_grad = np.stack([G.through_log_softmax(softmax_of[b], want_grad[i]) for i, b in enumerate(idx)])
    got_grad = grad[idx].astype(np.float64)
    scale = np.abs(want_grad).max()
    err = np.abs(got_grad - want_grad)
    tol = 1e-4 * np.abs(want_grad) + 1e-4 * scale
    worst = float((err / tol).max())
    nfall = int((flags[idx] != 0).sum())
    record(name, utterances=len(idx), fallback=nfall, max_rel_loss_err=float(rel.max()),
           max_abs_grad_err=float(err.max()), grad_max_norm=float(scale),
           max_grad_err_over_tol=worst, max_grad_err_rel_to_max_norm=float(err.max() / scale))
    assert nfall <= max_fallback, "%d of %d compared utterances were recomputed by the fallback kernel" % (nfall, len(idx))
    assert rel.max() <= 1e-4, "loss rel err %.3e" % rel.max()
    assert worst <= 1.0, "gradient error %.3f x tolerance (max abs %.3e, max-norm %.3e)" % (worst, err.max(), scale)


def synth_cfg(B, T, C, L, seed, ragged=None):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, T, C, generator=g)
    if ragged is None:
        tg = torch.randint(C - 2, (B, L), generator=g).tolist()   # randint(C-2), blank = C-1 (ctc_benchmark.py:23)
    else:
        lens = torch.randint(ragged[0], ragged[1] + 1, (B,), generator=g).tolist()
        tg = [torch.randint(C - 2, (n,), generator=g).tolist() for n in lens]
    return x, tg


# ------------------------------------------------------------------------------ cfg2
@pytest.mark.parametrize("kind", ["log_softmax", "raw"])
def test_cfg2_full_batch_elementwise(kind):
    """BASELINE configs[1]: the bench kernel at the bench shape, all 256 utterances."""
    B, T, C, L = 256, 1000, 30, 176
    x, tg = synth_cfg(B, T, C, L, 0)
    e = torch.log_softmax(x, 2) if kind == "log_softmax" else x
    losses, mean, grad, flags = ctc_capi(e.cuda(), tg, C - 1)
    compare("cfg2_" + kind, losses, grad, flags, e.numpy(), tg, C - 1, list(range(B)), [1.0 / B] * B)
    assert abs(mean - losses.mean()) <= 1e-5 * abs(losses.mean())


def test_cfg2_through_the_function_matches_the_abi():
    """CTCLoss(lp, list_of_lists, blank).backward() (benchmarks/ctc_benchmark.py:23-29) returns
    the ABI's numbers."""
    from gtn_applications_b200.criterions.ctc import CTCLoss
    B, T, C, L = 256, 1000, 30, 176
    x, tg = synth_cfg(B, T, C, L, 0)
    e = torch.log_softmax(x, 2).cuda()
    losses, mean, grad, _ = ctc_capi(e, tg, C - 1)
    lp = e.clone().requires_grad_(True)
    loss = CTCLoss(lp, tg, C - 1)
    loss.backward()
    assert abs(loss.item() - mean) <= 1e-6 * abs(mean)
    got = lp.grad.cpu().numpy()
    diff = float(np.abs(got - grad).max())
    record("cfg2_function_vs_abi", max_abs_diff=diff, bit_identical=bool(np.array_equal(got, grad)))
    assert diff <= 1e-6 * float(np.abs(grad).max())


def test_cfg2_fused_logits_elementwise():
    B, T, C, L = 256, 1000, 30, 176
    x, tg = synth_cfg(B, T, C, L, 1)
    losses, mean, grad, flags = ctc_capi(x.cuda(), tg, C - 1, logits=True)
    idx = list(range(0, B, 4))
    compare("cfg2_fused_logits", losses, grad, flags, G.log_softmax(x.numpy()), tg, C - 1, idx, [1.0 / B] * B,
            softmax_of=x.numpy())


def test_cfg2_ragged_targets_elementwise():
    """SURVEY §8(d) secondary variant: L_b ~ U{100..250} (pairs of unequal length in a block)."""
    B, T, C = 256, 1000, 30
    x, tg = synth_cfg(B, T, C, None, 2, ragged=(100, 250))
    e = torch.log_softmax(x, 2)
    scales = [1.0 / (len(t) * B) for t in tg]     # reduction="mean"
    losses, mean, grad, flags = ctc_capi(e.cuda(), tg, C - 1, scales=scales)
    idx = list(range(0, B, 2))
    compare("cfg2_ragged", losses, grad, flags, e.numpy(), tg, C - 1, idx, scales)


@pytest.mark.parametrize("scale", [2.0, 3.0])
def test_cfg2_steep_emissions_elementwise(scale):
    """log_softmax(s * randn) with targets unrelated to the scores: whatever kernel ends up
    computing an utterance, the result holds the tolerance (the fallback rate is recorded)."""
    B, T, C, L = 64, 1000, 30, 176
    x, tg = synth_cfg(B, T, C, L, 3)
    e = torch.log_softmax(x * scale, 2)
    losses, mean, grad, flags = ctc_capi(e.cuda(), tg, C - 1)
    record("cfg2_steep_%g_fallback" % scale, fallback_rate=float((flags != 0).mean()))
    compare("cfg2_steep_%g" % scale, losses, grad, flags, e.numpy(), tg, C - 1, list(range(0, B, 4)), [1.0 / B] * B,
            max_fallback=B)


# ------------------------------------------------------------------------------ cfg5
def test_cfg5_shard_slice_elementwise():
    """BASELINE configs[4] per-GPU shard (B=256, T=1500, C=80, L=264)."""
    B, T, C, L = 256, 1500, 80, 264
    x, tg = synth_cfg(B, T, C, L, 5)
    e = torch.log_softmax(x, 2)
    losses, mean, grad, flags = ctc_capi(e.cuda(), tg, C - 1)
    idx = list(range(0, B, 8))
    compare("cfg5_shard", losses, grad, flags, e.numpy(), tg, C - 1, idx, [1.0 / B] * B)
    rows = grad.sum(2)
    np.testing.assert_allclose(rows, np.full_like(rows, -1.0 / B), rtol=2e-4)


# ------------------------------------------------------------------------------ cfg4
def reference_word_pieces():
    """transducer_benchmark.py:19-23"""
    with open(os.path.join(G.GOLDEN, "word_pieces_tokens_1000.txt"), "r") as fid:
        tokens = sorted([l.strip() for l in fid])
    graphemes = sorted(set(c for t in tokens for c in t))
    return tokens, {t: i for i, t in enumerate(graphemes)}


def test_cfg4_reference_token_list_against_oracle(gtn64):
    """BASELINE configs[3] on the reference's token list: alignment graph indices bit-exact with the
    oracle's GTN restatement (transducer.py:265-276), loss + gradient of an utterance against the
    float64 DP over that acceptor, at T=1000, C=1001 (1000 pieces + blank), 150 pieces per target."""
    import dp_numpy
    import ref_criterions as rc
    from gtn_applications_b200.criterions.transducer import Transducer, TransducerLoss
    tokens, g2i = reference_word_pieces()
    assert len(tokens) == 1000
    rnd = random.Random(0)
    B, T, NP = 2, 1000, 150
    C = len(tokens) + 1
    crit = Transducer(tokens, g2i, blank="optional", allow_repeats=False, reduction="mean")
    targets = [[g2i[l] for wp in (rnd.choice(tokens) for _ in range(NP)) for l in wp] for _ in range(B)]
    torch.manual_seed(0)
    x = torch.randn(B, T, C)
    e = torch.log_softmax(x, 2)
    # oracle graphs (the reference's constructors on the oracle's GTN restatement) and the
    # product's host library, one alignment acceptor per utterance
    import ctypes
    from gtn_applications_b200 import _lib, graph as PG
    o = rc.Transducer(gtn64, tokens, g2i, blank="optional", allow_repeats=False)
    o.tokens.arc_sort(True)
    crit.tokens.arc_sort(True)
    flat = np.array([t for y in targets for t in y], dtype=np.int32)
    off = np.zeros(B + 1, dtype=np.int32)
    off[1:] = np.cumsum([len(y) for y in targets])
    hs = (ctypes.c_int32 * B)()
    _lib.check(_lib.lib().wfst_transducer_alignment_graphs(crit.tokens._h, crit.lexicon._h, flat.ctypes.data,
                                                           off.ctypes.data, B, hs))
    losses = []
    grads = []
    for b in range(B):
        oa = rc.graph_arrays(o.alignment_graph(targets[b]))
        pa = PG.Graph(_handle=hs[b]).arrays()
        for k in ("start", "accept", "src", "dst", "ilabel", "olabel"):
            assert np.array_equal(oa[k], pa[k]), "alignment graph %s differs from the oracle" % k
        Z, gE, _ = dp_numpy.acceptor_forward_backward(e[b].numpy(), oa["start"], oa["accept"], oa["src"], oa["dst"],
                                                      oa["ilabel"], oa["weight"])
        losses.append(-Z / len(targets[b]))
        grads.append(-gE / (len(targets[b]) * B))
    ed = e.cuda().requires_grad_(True)
    loss = TransducerLoss(ed, targets, crit.tokens, crit.lexicon, None, None, "mean")
    loss.backward()
    want = float(np.mean(losses))
    got_grad = ed.grad.cpu().numpy().astype(np.float64)
    want_grad = np.stack(grads)
    scale = np.abs(want_grad).max()
    err = np.abs(got_grad - want_grad)
    worst = float((err / (1e-4 * np.abs(want_grad) + 1e-4 * scale)).max())
    record("cfg4_reference_tokens", utterances=B, rel_loss_err=abs(loss.item() - want) / abs(want),
           max_abs_grad_err=float(err.max()), grad_max_norm=float(scale), max_grad_err_over_tol=worst)
    assert abs(loss.item() - want) <= 1e-4 * abs(want)
    assert worst <= 1.0
