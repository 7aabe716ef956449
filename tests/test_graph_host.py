"""CPU tests of the product's host-side graph library (csrc/graph.cpp through the C ABI;
no GPU needed): node / arc numbering and arc-list order must be BIT-EXACT with the oracle
restatement of GTN and with the fixtures produced from the reference's own constructors
("arc/state indices bit-exact", north_star)."""
import os

import numpy as np
import pytest

import _golden as GOLD
import ref_criterions as rc
from gtn_applications_b200 import graph as G
from gtn_applications_b200.criterions import asg as pasg, ctc as pctc, stc as pstc, transducer as ptr

KEYS = ("start", "accept", "src", "dst", "ilabel", "olabel")


def same(prod, oracle_graph, weights=True):
    x, y = prod.arrays(), rc.graph_arrays(oracle_graph)
    for k in KEYS:
        np.testing.assert_array_equal(x[k], y[k], err_msg=k)
    if weights:
        np.testing.assert_allclose(x["weight"], y["weight"], rtol=1e-6)
    np.testing.assert_array_equal(prod.arc_order(False), np.array(oracle_graph.out_order(), dtype=np.int32))
    np.testing.assert_array_equal(prod.arc_order(True), np.array(oracle_graph.in_order(), dtype=np.int32))


def test_ctc_asg_stc_constructors_match_oracle_and_fixtures(gtn32):
    z = GOLD.load("ctc")
    g = pctc.CTCLossFunction.create_ctc_graph([3, 3, 1, 0, 0, 2], 5)
    same(g, rc.ctc_graph(gtn32, [3, 3, 1, 0, 0, 2], 5))
    ref = GOLD.graph_of(z, "graph")
    for k in KEYS:
        np.testing.assert_array_equal(g.arrays()[k], ref[k])
    np.testing.assert_array_equal(g.arc_order(True), z["graph_in_order"])
    np.testing.assert_array_equal(g.arc_order(False), z["graph_out_order"])
    import torch
    tr = torch.randn(4, 3)
    same(pasg.ASGLossFunction.create_transitions_graph(tr), rc.asg_transitions_graph(gtn32, tr.numpy()))
    same(pasg.ASGLossFunction.create_force_align_graph([2, 0, 0, 1]), rc.asg_force_align_graph(gtn32, [2, 0, 0, 1]))
    same(pstc.STCLossFunction.create_stc_graph([2, 1, 1], 4, 0.5), rc.stc_graph(gtn32, [2, 1, 1], 4, 0.5))
    z = GOLD.load("stc")
    ref = GOLD.graph_of(z, "graph")
    for k in KEYS:
        np.testing.assert_array_equal(pstc.STCLossFunction.create_stc_graph([2, 1, 1], 4, 0.5).arrays()[k], ref[k])


@pytest.mark.parametrize("blank,rep", [("none", True), ("optional", True), ("optional", False), ("forced", True)])
def test_transducer_graphs_match_oracle(gtn32, blank, rep):
    import ctypes
    from gtn_applications_b200 import _lib
    tokens, g2i = ["a", "b", "ab", "ba", "aba"], {"a": 0, "b": 1}
    o = rc.Transducer(gtn32, tokens, g2i, blank=blank, allow_repeats=rep)
    tk = ptr.make_token_graph(tokens, blank, rep)
    lx = ptr.make_lexicon_graph(tokens, g2i)
    same(tk, o.tokens)
    same(lx, o.lexicon)
    tg = [[0, 1, 0], [1, 1, 0, 1, 0, 0], [0], [1, 0, 0, 0, 1, 1, 0, 1]]
    flat = np.array([t for x in tg for t in x], dtype=np.int32)
    off = np.zeros(len(tg) + 1, dtype=np.int32)
    off[1:] = np.cumsum([len(t) for t in tg])
    hs = (ctypes.c_int32 * len(tg))()
    _lib.check(_lib.lib().wfst_transducer_alignment_graphs(tk._h, lx._h, flat.ctypes.data,
                                                           off.ctypes.data, len(tg), hs))
    o.tokens.arc_sort(True)
    for b, y in enumerate(tg):
        same(G.Graph(_handle=hs[b]), o.alignment_graph(y))


def test_wordpiece_graphs_match_fixtures():
    z = GOLD.load("transducer")
    tokens, g2i = ["a", "b", "ab", "ba", "aba"], {"a": 0, "b": 1}
    for name, blank, rep in (("wp_none", "none", True), ("wp_opt", "optional", True),
                             ("wp_norep", "optional", False), ("wp_forced", "forced", True)):
        for prefix, g in (("_tokens", ptr.make_token_graph(tokens, blank, rep)),
                          ("_lexicon", ptr.make_lexicon_graph(tokens, g2i))):
            ref = GOLD.graph_of(z, name + prefix)
            for k in KEYS:
                np.testing.assert_array_equal(g.arrays()[k], ref[k], err_msg=name + prefix + k)


@pytest.mark.parametrize("ngram", [1, 2, 3])
def test_ngram_transition_graphs(gtn32, ngram):
    same(ptr.make_transitions_graph(ngram, 4), rc.ngram_transitions_graph(gtn32, ngram, 4))


def test_compose_with_epsilon_transitions_and_provenance(gtn32, tmp_path):
    # loaded back-off graph (tests/trans_backoff_test.txt content comes from the fixture)
    z = GOLD.load("transducer")
    a = GOLD.graph_of(z, "backoff_file")
    g = G.Graph(True)
    for s, acc in zip(a["start"], a["accept"]):
        g.add_node(bool(s), bool(acc))
    g.add_arcs(a["src"], a["dst"], a["ilabel"], a["olabel"], a["weight"])
    og = gtn32.Graph(True)
    for s, acc in zip(a["start"], a["accept"]):
        og.add_node(bool(s), bool(acc))
    for s, d, i, o, w in zip(a["src"], a["dst"], a["ilabel"], a["olabel"], a["weight"]):
        og.add_arc(int(s), int(d), int(i), int(o), float(w))
    g.arc_sort()
    og.arc_sort()
    N = 5
    tokens = [(n,) for n in range(N)]
    g2i = {n: n for n in range(N)}
    o = rc.Transducer(gtn32, tokens, g2i, blank="optional", allow_repeats=False)
    o.tokens.arc_sort(True)
    oal = o.alignment_graph([0, 1, 0])
    tk = ptr.make_token_graph(tokens, "optional", False)
    lx = ptr.make_lexicon_graph(tokens, g2i)
    tk.arc_sort(True)
    tgt = ptr.make_chain_graph([0, 1, 0])
    tgt.arc_sort(True)
    dec = G.remove(G.project_output(G.compose(tgt, lx)))
    dec.arc_sort()
    pal = G.project_input(G.remove(G.compose(tk, dec)))
    pal.arc_sort()
    same(pal, oal)
    same(G.intersect(g, pal), gtn32.intersect(og, oal))
    # emissions o transitions with epsilon back-off arcs (transducer.py:287)
    same(G.intersect(G.linear_graph(3, N + 1), g), gtn32.intersect(gtn32.linear_graph(3, N + 1), og))
    # text / binary round trips
    p = os.path.join(tmp_path, "g.txt")
    G.savetxt(p, g)
    assert G.equal(G.loadtxt(p), g)
    same(G.loadtxt(p), gtn32.loadtxt(p))
    p = os.path.join(tmp_path, "g.bin")
    G.save(p, g)
    assert G.equal(G.load(p), g)
    gtn32.save(os.path.join(tmp_path, "o.bin"), og)
    assert G.equal(G.load(os.path.join(tmp_path, "o.bin")), g)   # same binary format as the oracle's
    # layout (as recalled from GTN's saveGraph): num_arcs follows the start / accept id lists
    raw = np.fromfile(p, dtype=np.int32)
    a = g.arrays()
    ns, na = int(a["start"].sum()), int(a["accept"].sum())
    assert raw[0] == g.num_nodes() and raw[1] == ns and raw[2] == na and raw[3 + ns + na] == g.num_arcs()
    # files written by round 1 of this library (num_arcs as the 4th header word) still load
    old = np.concatenate((raw[:3], [g.num_arcs()], raw[3:3 + ns + na], raw[4 + ns + na:])).astype(np.int32)
    p1 = os.path.join(tmp_path, "round1.bin")
    old.tofile(p1)
    assert G.equal(G.load(p1), g)
    # corrupt headers are rejected with an error (nothing may throw across the C boundary)
    for bad in (np.array([-5, 1, 1, 0], np.int32), np.array([3, -1, 0, 0], np.int32),
                np.array([2, 1, 1, 0, 1, 7, 0, 0, 0, 0, 0], np.int32), raw[:-2], np.array([1], np.int32)):
        pb = os.path.join(tmp_path, "bad.bin")
        bad.tofile(pb)
        with pytest.raises((ValueError, RuntimeError)):
            G.load(pb)


def test_isomorphic_equal_and_viterbi_path(gtn32):
    a = ptr.make_transitions_graph(2, 3)
    b = ptr.make_transitions_graph(2, 3)
    assert G.equal(a, b) and G.isomorphic(a, b)
    c = ptr.make_transitions_graph(2, 4)
    assert not G.isomorphic(a, c)
    # best path with ties: first maximum in in-list order, as the oracle does
    rng = np.random.default_rng(0)
    for trial in range(20):
        g, og = G.Graph(False), gtn32.Graph(False)
        n = 8
        for k in range(n):
            g.add_node(k == 0, k == n - 1)
            og.add_node(k == 0, k == n - 1)
        for k in range(30):
            s = int(rng.integers(0, n - 1))
            d = int(rng.integers(s + 1, n))
            lab, w = int(rng.integers(0, 5)), float(rng.integers(0, 3))
            g.add_arc(s, d, lab, lab + 1, w)
            og.add_arc(s, d, lab, lab + 1, w)
        same(G.viterbi_path(g), gtn32.viterbi_path(og))


def test_errors():
    g = G.Graph()
    g.add_node()
    with pytest.raises(ValueError):
        g.add_arc(0, 3, 1)
    with pytest.raises(ValueError):
        ptr.make_token_graph(["a"], blank="none", allow_repeats=False)
    with pytest.raises(ValueError):
        ptr.Transducer(["a"], {"a": 0}, blank="sometimes")
    with pytest.raises(ValueError):
        ptr.Transducer(["a"], {"a": 0}, ngram=1, transitions=G.Graph())


def test_conv_kernel_graphs_match_reference_constructor():
    """make_kernel_graph (criterions/transducer.py:351-367): node flags and arc lists
    bit-exact with what the reference's constructor produced on the oracle shim."""
    import _golden as Gd
    from gtn_applications_b200.criterions.transducer import make_kernel_graph
    z = Gd.load("conv")
    for i in range(5):
        tok = z["kg%d_token" % i].tolist()
        opt, spike = (bool(v) for v in z["kg%d_flags" % i])
        g = make_kernel_graph(tok, 2, opt, spike=spike).arrays()
        for k in ("start", "accept", "src", "dst", "ilabel", "olabel"):
            assert np.array_equal(np.asarray(g[k]).astype(np.int64), z["kg%d_%s" % (i, k)].astype(np.int64)), (i, k)


@pytest.mark.parametrize("blank,rep", [("none", True), ("optional", True), ("optional", False), ("forced", True)])
def test_batched_decode_matches_the_per_utterance_pipeline(blank, rep):
    """wfst_transducer_decode_paths (alignment -> tokens for a whole batch on host threads) against
    the reference's per-utterance sequence compose(chain, tokens) -> viterbi_path -> project_output
    -> remove (criterions/transducer.py:223-233) run through the same host library one call at a time."""
    import ctypes
    import random
    from gtn_applications_b200 import _lib
    rnd = random.Random(5)
    pieces = ["a", "b", "ab", "ba", "aba", "c", "ca"]
    tokens = ptr.make_token_graph(pieces, blank=blank, allow_repeats=rep)
    tokens.arc_sort()
    V = len(pieces) + (0 if blank == "none" else 1)
    B, T = 9, 23
    labels = np.array([[rnd.randrange(V) for _ in range(T)] for _ in range(B)], dtype=np.int32)
    labels[0, :] = labels[0, 0]                     # one alignment that is a single repeated token
    out = np.zeros((B, T), dtype=np.int32)
    counts = np.zeros(B, dtype=np.int32)
    _lib.check(_lib.lib().wfst_transducer_decode_paths(
        tokens._h, labels.ctypes.data, B, T, out.ctypes.data, counts.ctypes.data))
    for b in range(B):
        chain = G.Graph(False)
        chain.add_node(True, False)
        for i, lab in enumerate(labels[b].tolist()):
            chain.add_node(False, i == T - 1)
            chain.add_arc(i, i + 1, int(lab))
        want = G.remove(G.project_output(G.viterbi_path(G.compose(chain, tokens)))).labels_to_list()
        assert out[b, :counts[b]].tolist() == want


def test_alignment_graph_cache_returns_the_same_graphs_and_freezes_them():
    """wfst_transducer_alignment_cache: a target met again comes from the LRU (same arrays, bit
    for bit, as a freshly built graph); cached graphs reject modification through a handle."""
    import ctypes
    from gtn_applications_b200 import _lib, graph as PG
    from gtn_applications_b200.criterions.transducer import make_lexicon_graph, make_token_graph
    L = _lib.lib()
    tokens = ["a", "b", "ab", "ba", "aba", "bb"]
    g2i = {"a": 0, "b": 1}
    tk = make_token_graph(tokens, blank="optional", allow_repeats=False)
    lx = make_lexicon_graph(tokens, g2i)
    targets = [[0, 1, 0], [1, 1, 0, 0, 1], [0], [0, 1, 0]]
    flat = np.array([t for y in targets for t in y], dtype=np.int32)
    off = np.zeros(len(targets) + 1, dtype=np.int32)
    off[1:] = np.cumsum([len(y) for y in targets])
    B = len(targets)

    def build():
        hs = (ctypes.c_int32 * B)()
        _lib.check(L.wfst_transducer_alignment_graphs(tk._h, lx._h, flat.ctypes.data, off.ctypes.data, B, hs))
        return [PG.Graph(_handle=h) for h in hs]

    hits, misses = ctypes.c_ulonglong(), ctypes.c_ulonglong()
    _lib.check(L.wfst_transducer_alignment_cache(0, None, None))          # off
    plain = [g.arrays() for g in build()]
    _lib.check(L.wfst_transducer_alignment_cache(1 << 20, None, None))    # on, empty
    first = build()
    second = build()
    _lib.check(L.wfst_transducer_alignment_cache(-1, ctypes.byref(hits), ctypes.byref(misses)))
    assert misses.value >= 3 and hits.value >= B          # utterances 0 and 3 share a target
    for ref, a, b in zip(plain, first, second):
        for k in ("start", "accept", "src", "dst", "ilabel", "olabel", "weight"):
            assert np.array_equal(ref[k], a.arrays()[k]) and np.array_equal(ref[k], b.arrays()[k])
    with pytest.raises(Exception):
        second[0].add_node()
    second[0].arc_sort()          # already ilabel-sorted: a no-op, allowed
    _lib.check(L.wfst_transducer_alignment_cache(16 << 20, None, None))   # back to the default capacity


def test_stc_graphs_of_a_batch_equal_the_per_utterance_builder():
    """wfst_stc_graphs (the whole batch in the host library) against STCLossFunction.create_stc_graph
    (criterions/stc.py:22-64, itself checked against the oracle) + arc_sort: arrays bit-exact,
    incl. an empty target, repeated labels and a single label; labels out of range are refused."""
    import ctypes
    import math
    from gtn_applications_b200 import _lib, graph as G
    from gtn_applications_b200.criterions.stc import STCLossFunction, STC_BLANK_IDX
    L = _lib.lib()
    star, prob = 7, 0.37
    targets = [[1, 2, 2, 5, 1], [], [3], [6, 6, 6, 6], [1, 2, 3, 4, 5, 6, 1, 2, 3]]
    flat = np.array([t for y in targets for t in y], dtype=np.int32)
    offs = np.zeros(len(targets) + 1, dtype=np.int32)
    offs[1:] = np.cumsum([len(y) for y in targets])
    hs = (ctypes.c_int32 * len(targets))()
    _lib.check(L.wfst_stc_graphs(flat.ctypes.data, offs.ctypes.data, len(targets), star, math.log(prob),
                                 STC_BLANK_IDX, hs))
    for h, y in zip(hs, targets):
        got = G.Graph(_handle=h)
        want = STCLossFunction.create_stc_graph(y, star, prob)
        want.arc_sort(False)
        ga, wa = got.arrays(), want.arrays()
        for k in wa:
            assert np.array_equal(ga[k], wa[k]), k
        assert np.array_equal(got.arc_order(False), want.arc_order(False))
        assert np.array_equal(got.arc_order(True), want.arc_order(True))
    bad = np.array([1, 7], dtype=np.int32)
    with pytest.raises(ValueError):
        _lib.check(L.wfst_stc_graphs(bad.ctypes.data, np.array([0, 2], dtype=np.int32).ctypes.data, 1, star,
                                     math.log(prob), STC_BLANK_IDX, hs))
