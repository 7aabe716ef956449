"""GPU parity for the criteria that run on the generic lattice kernel: STC
(STCLossFunction / STC module) and the word-piece transducer (TransducerLossFunction /
Transducer module, with and without an epsilon-free transition graph), plus the Viterbi
decoders, against the float64 oracle, the reference's known answers and the committed
fixtures.  Tolerance: see test_gpu_ctc.py."""
import math

import numpy as np
import pytest
import torch

import _golden as G
import ref_criterions as rc
from test_gpu_ctc import assert_close, assert_close_f32_fixture

pytestmark = pytest.mark.gpu


# ------------------------------------------------------------------------- STC
def test_stc_known_answers():
    from gtn_applications_b200.criterions.stc import STC
    with np.errstate(divide="ignore"):
        lp = torch.log(torch.tensor([0.0, 1.0, 1.0, 0.0, 0.0, 1.0]).view(3, 1, 2)).cuda()
    assert abs(STC(0, 1, 1, 1)(lp, [[1, 1]]).item()) < 1e-6               # gtn_stc_test.py:25-37
    lp = torch.log_softmax(torch.zeros(3, 1, 4), 2).cuda()
    loss = STC(0, 1, 1, 1, "none")(lp, [[1, 2]])
    assert abs(loss.item() + math.log(0.25 * 0.25 * 2.5)) < 1e-5          # gtn_stc_test.py:39-51


@pytest.mark.parametrize("case", ["fn_none", "fn_mean"])
def test_stc_function_fixtures_and_oracle(gtn64, case, lattice_kernel):
    from gtn_applications_b200.criterions.stc import STCLoss
    z = G.load("stc")
    tg = G.unpack(z[case + "_targets"], z[case + "_offsets"])
    x = torch.tensor(z[case + "_emissions"], device="cuda", requires_grad=True)
    loss = STCLoss(x, tg, float(z[case + "_prob"]), str(z[case + "_reduction"]))
    loss.backward()
    want = float(z[case + "_loss"])
    assert abs(loss.item() - want) <= 1e-4 * max(1.0, abs(want))
    assert_close_f32_fixture(x.grad.cpu().numpy(), z[case + "_grad"])
    ref = rc.stc(gtn64, z[case + "_emissions"], tg, float(z[case + "_prob"]), str(z[case + "_reduction"]))
    assert abs(loss.item() - ref["loss"]) <= 1e-4 * abs(ref["loss"])
    assert_close(x.grad.cpu().numpy(), ref["grad"])


def test_stc_function_at_a_training_shape_against_the_float64_oracle(gtn64, lattice_kernel):
    """B=4, T=300, 15 tokens (C* = 32 with the <star> columns), L up to 40 -- several emission tiles,
    acceptors of ~200 nodes from wfst_stc_graphs -- loss and gradient against the gtn64 oracle."""
    from gtn_applications_b200.criterions.stc import STCLoss
    rng = np.random.default_rng(21)
    B, T, star = 4, 300, 16
    tg = [rng.integers(1, star, size=n).tolist() for n in (40, 1, 23, 31)]
    em = np.log(rng.dirichlet(np.ones(2 * star), size=(B, T))).astype(np.float32)
    x = torch.tensor(em, device="cuda", requires_grad=True)
    loss = STCLoss(x, tg, 0.3, "mean")
    loss.backward()
    ref = rc.stc(gtn64, em, tg, 0.3, "mean")
    assert abs(loss.item() - ref["loss"]) <= 1e-4 * abs(ref["loss"])
    assert_close(x.grad.cpu().numpy(), ref["grad"])


def test_stc_module_fixture():
    from gtn_applications_b200.criterions.stc import STC
    z = G.load("stc")
    crit = STC(0, p0=0.5, plast=0.9, thalf=3, reduction="mean")
    crit.train()
    logits = torch.tensor(z["module_logits"], device="cuda", requires_grad=True)
    tg = G.unpack(z["module_targets"], z["module_offsets"])
    loss = crit(torch.log_softmax(logits, 2), tg)
    loss.backward()
    assert crit.nstep == 1
    assert abs(loss.item() - float(z["module_loss"])) <= 1e-4 * abs(float(z["module_loss"]))
    assert_close_f32_fixture(logits.grad.cpu().numpy(), z["module_grad_logits"])


def test_stc_errors():
    from gtn_applications_b200.criterions.stc import STC, STCLoss
    with pytest.raises(ValueError, match="invalid value for reduction"):
        STCLoss(torch.zeros(1, 3, 4, device="cuda"), [[1]], 0.5, "sum")
    with pytest.raises(AssertionError):
        STC(1)


# ------------------------------------------------------------------ transducer
TOKENS, G2I = ["a", "b", "ab", "ba", "aba"], {"a": 0, "b": 1}


@pytest.mark.parametrize("name,blank,rep", [("wp_none", "none", True), ("wp_opt", "optional", True),
                                            ("wp_norep", "optional", False), ("wp_forced", "forced", True)])
def test_transducer_wordpiece_fixtures(name, blank, rep, lattice_kernel):
    from gtn_applications_b200.criterions.transducer import Transducer
    z = G.load("transducer")
    tg = G.unpack(z["wp_targets"], z["wp_offsets"])
    crit = Transducer(TOKENS, G2I, blank=blank, allow_repeats=rep, reduction="mean")
    x = torch.tensor(z[name + "_logits"], device="cuda", requires_grad=True)
    loss = crit(x, tg)
    loss.backward()
    want = float(z[name + "_loss"])
    assert abs(loss.item() - want) <= 1e-4 * max(1.0, abs(want))
    assert_close_f32_fixture(x.grad.cpu().numpy(), z[name + "_grad_logits"])
    pred = crit.viterbi(x.detach())
    assert [p.tolist() for p in pred] == G.unpack(z[name + "_viterbi"], z[name + "_viterbi_offsets"])


def test_transducer_known_answers():
    from gtn_applications_b200.criterions.transducer import Transducer
    import test_oracle_golden as lit
    lp = torch.log(torch.tensor([1.0, 0.0, 0.0, 1.0, 1.0, 0.0]).view(1, 3, 2)).cuda()
    # transducer_test.py:100-126 (no blank / optional blank / no repeats)
    assert abs(Transducer(["a", "b"], {"a": 0, "b": 1})(lp, [[0, 1, 0]]).item()) < 1e-6
    assert abs(Transducer(["a"], {"a": 0}, blank="optional")(lp, [[0, 0]]).item()) < 1e-6
    assert abs(Transducer(["a"], {"a": 0}, blank="optional", allow_repeats=False)(lp, [[0, 0]]).item()) < 1e-6
    # transducer_test.py:143-216: the two warp-ctc vectors through the transducer
    toks, g2i = ["a", "b", "c", "d", "e"], {"a": 0, "b": 1, "c": 2, "d": 3, "e": 4}
    for probs, labels, want, want_grad, rep in ((lit.WARP_CTC_1, [[0, 1, 2, 1, 0]], 3.34211, lit.WARP_CTC_1_GRAD, True),
                                                (lit.WARP_CTC_2, [[0, 1, 1, 0]], 5.42262, lit.WARP_CTC_2_GRAD, False)):
        x = torch.log(torch.tensor(probs, dtype=torch.float32)).cuda().requires_grad_(True)
        loss = Transducer(toks, g2i, blank="optional", allow_repeats=rep)(x, labels)
        loss.backward()
        assert abs(loss.item() - want) < 5e-5
        np.testing.assert_allclose(x.grad.cpu().numpy(), want_grad, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("reduction", ["none", "mean"])
def test_transducer_equals_ctc(reduction):
    # transducer_test.py:275-316: single-grapheme tokens + optional blank + no repeats = CTC
    from gtn_applications_b200.criterions.transducer import Transducer
    from gtn_applications_b200.criterions.ctc import CTCLoss
    T, N, B = 20, 15, 5
    tgt = [[0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10], [1, 1], [0, 2, 3], [0, 0, 0, 0, 0], [0, 4, 8, 12]]
    x = torch.randn(B, T, N, generator=torch.Generator().manual_seed(0)).cuda().requires_grad_(True)
    crit = Transducer([(t,) for t in range(N - 1)], {t: t for t in range(N - 1)}, blank="optional",
                      allow_repeats=False, reduction=reduction)
    a = CTCLoss(torch.log_softmax(x, 2), tgt, N - 1, reduction)
    a.backward()
    ga = x.grad.clone()
    x.grad = None
    b = crit(x, tgt)
    b.backward()
    assert abs(a.item() - b.item()) <= 1e-4 * abs(a.item())
    torch.testing.assert_close(ga, x.grad, rtol=1e-4, atol=1e-5)


def test_transducer_with_asg_transitions(gtn64, lattice_kernel):
    # transducer_test.py:420-508: ASG as a transducer with a learned bigram transition graph
    from gtn_applications_b200.criterions.transducer import Transducer
    from gtn_applications_b200.criterions.asg import ASGLossFunction
    import test_oracle_golden as lit
    N = 6
    crit = Transducer([(n,) for n in range(N)], {n: n for n in range(N)},
                      transitions=ASGLossFunction.create_transitions_graph(torch.zeros(N + 1, N))).cuda()
    x = torch.tensor(lit.ASG_EMISSIONS, dtype=torch.float32, device="cuda", requires_grad=True)
    loss = crit(x, lit.ASG_LABELS)
    loss.backward()
    assert abs(loss.item() - 7.47995) < 5e-5
    np.testing.assert_allclose(x.grad.cpu().numpy(), lit.ASG_GRAD, rtol=1e-3, atol=1e-5)
    tg = crit.transition_params.grad[N:].view(N, N).cpu().numpy()
    np.testing.assert_allclose(tg, lit.ASG_TRANS_GRAD, rtol=1e-2, atol=1e-5)
    # random parameters against the float64 oracle
    rng = np.random.default_rng(1)
    params = rng.standard_normal(N + N * N).astype(np.float32) * 0.5
    crit.transition_params.data = torch.tensor(params, device="cuda")
    crit.transition_params.grad = None
    x.grad = None
    loss = crit(x, lit.ASG_LABELS)
    loss.backward()
    o = rc.Transducer(gtn64, [(n,) for n in range(N)], {n: n for n in range(N)},
                      transitions=rc.asg_transitions_graph(gtn64, np.zeros((N + 1, N), dtype=np.float32)))
    ref = o.loss(lit.ASG_EMISSIONS, lit.ASG_LABELS, params)
    assert abs(loss.item() - ref["loss"]) <= 1e-4 * abs(ref["loss"])
    assert_close(x.grad.cpu().numpy(), ref["grad"])
    assert_close(crit.transition_params.grad.cpu().numpy(), ref["grad_transitions"])
    assert list(crit.state_dict().keys()) == ["transition_params"]


def test_transducer_random_wordpieces_against_oracle(gtn64):
    from gtn_applications_b200.criterions.transducer import Transducer
    rng = np.random.default_rng(7)
    tokens = ["a", "b", "c", "ab", "bc", "ca", "abc", "cab", "aa"]
    g2i = {"a": 0, "b": 1, "c": 2}
    B, T = 4, 70
    tg = [rng.integers(0, 3, size=n).tolist() for n in (12, 1, 30, 7)]
    x = rng.standard_normal((B, T, len(tokens) + 1)).astype(np.float32)
    crit = Transducer(tokens, g2i, blank="optional", allow_repeats=False, reduction="mean")
    xt = torch.tensor(x, device="cuda", requires_grad=True)
    loss = crit(xt, tg)
    loss.backward()
    o = rc.Transducer(gtn64, tokens, g2i, blank="optional", allow_repeats=False, reduction="mean")
    ref = o.loss(G.log_softmax(x), tg)
    assert abs(loss.item() - ref["loss"]) <= 1e-4 * abs(ref["loss"])
    assert_close(xt.grad.cpu().numpy(), G.through_log_softmax(x, ref["grad"]))


@pytest.mark.parametrize("name,ngram,blank,rep", [("ngram1", 1, "optional", False), ("ngram2", 2, "optional", False),
                                                  ("ngram2_asg", 2, "none", True)])
def test_ngram_transitions_with_epsilon_arcs_match_reference_fixtures(name, ngram, blank, rep, lattice_kernel):
    """make_transitions_graph for ngram > 1 ends in epsilon </s> arcs (transducer.py:52-56):
    loss, emission gradient, transition-parameter gradient and the Viterbi decode against
    the reference's own module (fixtures from make_golden.py)."""
    from gtn_applications_b200.criterions.transducer import Transducer
    z = G.load("transducer")
    N = 4
    crit = Transducer([(i,) for i in range(N)], {i: i for i in range(N)}, ngram=ngram, blank=blank,
                      allow_repeats=rep, reduction="mean").cuda()
    crit.transition_params.data = torch.tensor(z[name + "_params"], device="cuda")
    x = torch.tensor(z[name + "_emissions"], device="cuda", requires_grad=True)
    tg = G.unpack(z[name + "_targets"], z[name + "_offsets"])
    loss = crit(x, tg)
    loss.backward()
    want = float(z[name + "_loss"])
    assert abs(loss.item() - want) <= 1e-4 * max(1.0, abs(want))
    assert_close_f32_fixture(x.grad.cpu().numpy(), z[name + "_grad"])
    assert_close_f32_fixture(crit.transition_params.grad.cpu().numpy(), z[name + "_grad_params"])
    vit = crit.viterbi(x.detach())
    assert [v.tolist() for v in vit] == G.unpack(z[name + "_viterbi"], z[name + "_viterbi_offsets"])


def test_backoff_transition_graph_with_epsilon_arcs_matches_reference_fixture(tmp_path):
    """A loaded back-off graph (tests/trans_backoff_test.txt, transducer_test.py:534-566): epsilon
    back-off arcs between n-gram states are folded into the arcs that follow them."""
    from gtn_applications_b200.criterions.transducer import Transducer
    from gtn_applications_b200 import graph as Gr
    z = G.load("transducer")
    g = Gr.Graph(False)
    for st, ac in zip(z["backoff_file_start"], z["backoff_file_accept"]):
        g.add_node(bool(st), bool(ac))
    for s_, d_, il, ol in zip(z["backoff_file_src"], z["backoff_file_dst"], z["backoff_file_ilabel"],
                              z["backoff_file_olabel"]):
        g.add_arc(int(s_), int(d_), int(il), int(ol), 0.0)
    N = 5
    crit = Transducer([(i,) for i in range(N)], {i: i for i in range(N)}, blank="optional", allow_repeats=False,
                      transitions=g).cuda()
    crit.transition_params.data = torch.tensor(z["backoff_params"], device="cuda")
    x = torch.tensor(z["backoff_emissions"], device="cuda", requires_grad=True)
    tg = G.unpack(z["backoff_targets"], z["backoff_offsets"])
    loss = crit(x, tg)
    loss.backward()
    want = float(z["backoff_loss"])
    assert abs(loss.item() - want) <= 1e-4 * max(1.0, abs(want))
    assert_close_f32_fixture(x.grad.cpu().numpy(), z["backoff_grad"])
    assert_close_f32_fixture(crit.transition_params.grad.cpu().numpy(), z["backoff_grad_params"])


# --------------------------------------------------------------------- viterbi
def test_asg_viterbi_known_answer_and_fixture():
    from gtn_applications_b200.criterions.asg import ASG
    # gtn_asg_test.py:107-124
    crit = ASG(3, 1, False).cuda()
    crit.transitions.data = torch.tensor([0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 0, 0, 0, 0, 2, 0, 2, 0, 0],
                                         dtype=torch.float32, device="cuda").view(5, 4)
    x = torch.tensor([0, 0, 0, 7, 0, 5, 4, 3, 0, 5, 8, 5, 0, 5, 4, 3], dtype=torch.float32, device="cuda").view(1, 4, 4)
    assert crit.viterbi(x)[0].tolist() == [2, 1, 0]
    z = G.load("asg")
    crit = ASG(4, num_replabels=2, use_garbage=True).cuda()
    crit.transitions.data = torch.tensor(z["module_transitions"], device="cuda")
    got = crit.viterbi(torch.tensor(z["module_emissions"], device="cuda"))
    assert [g.tolist() for g in got] == G.unpack(z["module_viterbi"], z["module_viterbi_offsets"])


@pytest.mark.parametrize("B,T,C,quant", [(7, 50, 30, False), (3, 1, 5, False), (5, 300, 32, True), (4, 64, 9, True)])
def test_asg_dense_viterbi_matches_generic_best_path(B, T, C, quant):
    """The dense warp-per-utterance best-path kernel against the generic kernel on the packed
    transition graph: same labels, also when scores tie (quantised scores force ties)."""
    from gtn_applications_b200 import graph as GG
    from gtn_applications_b200.criterions.asg import ASGLossFunction
    from gtn_applications_b200.decode import asg_viterbi_labels, lattice_viterbi
    rng = np.random.default_rng(B * 1000 + T + C)
    e = rng.standard_normal((B, T, C)).astype(np.float32)
    tr = rng.standard_normal((C + 1, C)).astype(np.float32)
    if quant:
        e, tr = np.round(e), np.round(tr)
    e, tr = torch.tensor(e, device="cuda"), torch.tensor(tr, device="cuda")
    from gtn_applications_b200 import _lib
    assert _lib.lib().wfst_asg_viterbi_supported(T, C) == 1
    dense = asg_viterbi_labels(e, tr)
    packed = GG.pack_graphs([ASGLossFunction.create_transitions_graph(tr)], e.device)
    scores, generic, _ = lattice_viterbi(e, packed, shared=True)
    assert torch.equal(dense, generic)
    # the path score recomputed from the labels equals the kernel's score
    lab = dense.long()
    path = e.gather(2, lab.unsqueeze(2)).squeeze(2).sum(1) + tr[0][lab[:, 0]]
    if T > 1:
        path = path + tr[1 + lab[:, 1:], lab[:, :-1]].sum(1)
    torch.testing.assert_close(path, scores, rtol=1e-5, atol=1e-4)


def test_transducer_viterbi_known_answers():
    # transducer_test.py:318-365 and 510-532
    from gtn_applications_b200.criterions.transducer import Transducer
    from gtn_applications_b200.criterions.asg import ASGLossFunction
    e1 = torch.tensor([0, 4, 0, 1, 0, 2, 1, 1, 0, 0, 0, 2, 0, 0, 0, 2, 8, 0, 0, 2], dtype=torch.float).view(5, 4)
    e2 = torch.tensor([0, 2, 1, 7, 0, 2, 9, 1, 0, 0, 0, 2, 0, 0, 5, 2, 1, 0, 0, 2], dtype=torch.float).view(5, 4)
    em = torch.stack([e1, e2]).cuda()
    crit = Transducer(["a", "b", "c", "d"], {"a": 0, "b": 1, "c": 2, "d": 3}, blank="none")
    assert [p.tolist() for p in crit.viterbi(em)] == [[1, 3, 0], [3, 2, 3, 2, 3]]
    crit = Transducer(["a", "b", "c"], {"a": 0, "b": 1, "c": 2}, blank="optional", allow_repeats=False)
    assert [p.tolist() for p in crit.viterbi(em)] == [[1, 0], [2, 2]]
    N = 3
    crit = Transducer([(n,) for n in range(N)], {n: n for n in range(N)},
                      transitions=ASGLossFunction.create_transitions_graph(torch.zeros(N + 1, N))).cuda()
    crit.transition_params.data = torch.tensor([0, 0, 0, 0, 2, 0, 0, 0, 2, 2, 0, 0], dtype=torch.float32, device="cuda")
    x = torch.tensor([0, 0, 7, 5, 4, 3, 5, 8, 5, 5, 4, 3], dtype=torch.float32, device="cuda").view(1, 4, 3)
    assert crit.viterbi(x)[0].tolist() == [2, 1, 0]


def test_full_size_properties_cfg4():
    """BASELINE configs[3] (transducer, 1000 word pieces, B=64, T=1000): size-independent checks.
    Every alignment crosses each frame once, so the emission-gradient rows sum to -scale_b / B;
    the gradient is non-positive; the shared-memory lattice kernel and the generic kernel agree;
    the alignment graphs of the batch equal the ones built one utterance at a time."""
    import random
    from gtn_applications_b200 import _lib
    from gtn_applications_b200.criterions.transducer import Transducer, TransducerLoss
    rnd = random.Random(0)
    letters = "abcdefghijklmnopqrstuvwxyz"
    pieces = sorted({"".join(rnd.choice(letters) for _ in range(rnd.randint(1, 4))) for _ in range(1400)})[:1000]
    pieces = sorted(set(pieces) | set(letters))
    g2i = {ch: i for i, ch in enumerate(letters)}
    B, T, NP = 64, 1000, 150
    crit = Transducer(pieces, g2i, blank="optional", allow_repeats=False, reduction="mean")
    C = len(pieces) + 1
    torch.manual_seed(0)
    e0 = torch.log_softmax(torch.randn(B, T, C, device="cuda"), 2)
    targets = [[g2i[c] for c in "".join(rnd.choice(pieces) for _ in range(NP))] for _ in range(B)]
    crit.tokens.arc_sort(True)

    def go():
        e = e0.clone().requires_grad_(True)
        loss = TransducerLoss(e, targets, crit.tokens, crit.lexicon, None, None, "mean")
        loss.backward()
        return loss.item(), e.grad

    loss, g = go()
    assert math.isfinite(loss) and loss > 0
    assert torch.all(g <= 1e-7)
    want = torch.tensor([-1.0 / (len(t) * B) for t in targets], device="cuda").unsqueeze(1).expand(B, T)
    torch.testing.assert_close(g.sum(2), want, rtol=3e-4, atol=0)
    old = _lib.lib().wfst_debug_force_generic_lattice(1)
    try:
        loss2, g2 = go()
    finally:
        _lib.lib().wfst_debug_force_generic_lattice(old)
    assert abs(loss - loss2) <= 1e-5 * abs(loss2)
    assert_close(g.cpu().numpy(), g2.cpu().numpy())
    # one utterance scored alone gives the same per-utterance gradient (x B)
    e = e0[:1].clone().requires_grad_(True)
    TransducerLoss(e, targets[:1], crit.tokens, crit.lexicon, None, None, "mean").backward()
    assert_close(e.grad[0].cpu().numpy() / B, g[0].cpu().numpy())
