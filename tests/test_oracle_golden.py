"""CPU tests (run everywhere): the oracle restatement that travels to the GPU box
(oracle/ref_criterions.py on the float32 and float64 `gtn` shims, and the
closed-form numpy DP oracle/dp_numpy.py) against
  * the reference's literal known-answer vectors (gtn_ctc_test.py:24-80,
    gtn_asg_test.py:25-124, gtn_stc_test.py:25-51, transducer_test.py:100-216),
  * the committed fixtures generated from the reference's own criterions
    (tests/golden/make_golden.py)."""
import math

import numpy as np
import pytest

import _golden as G
import dp_numpy
import ref_criterions as rc

# ---- literal vectors copied from the reference's tests (data, not code) ----
WARP_CTC_1 = np.array([
    0.633766, 0.221185, 0.0917319, 0.0129757, 0.0142857, 0.0260553,
    0.111121, 0.588392, 0.278779, 0.0055756, 0.00569609, 0.010436,
    0.0357786, 0.633813, 0.321418, 0.00249248, 0.00272882, 0.0037688,
    0.0663296, 0.643849, 0.280111, 0.00283995, 0.0035545, 0.00331533,
    0.458235, 0.396634, 0.123377, 0.00648837, 0.00903441, 0.00623107]).reshape(1, 5, 6)
WARP_CTC_1_GRAD = np.array([
    -0.366234, 0.221185, 0.0917319, 0.0129757, 0.0142857, 0.0260553,
    0.111121, -0.411608, 0.278779, 0.0055756, 0.00569609, 0.010436,
    0.0357786, 0.633813, -0.678582, 0.00249248, 0.00272882, 0.0037688,
    0.0663296, -0.356151, 0.280111, 0.00283995, 0.0035545, 0.00331533,
    -0.541765, 0.396634, 0.123377, 0.00648837, 0.00903441, 0.00623107]).reshape(1, 5, 6)
WARP_CTC_2 = np.array([
    0.30176, 0.28562, 0.0831517, 0.0862751, 0.0816851, 0.161508,
    0.24082, 0.397533, 0.0557226, 0.0546814, 0.0557528, 0.19549,
    0.230246, 0.450868, 0.0389607, 0.038309, 0.0391602, 0.202456,
    0.280884, 0.429522, 0.0326593, 0.0339046, 0.0326856, 0.190345,
    0.423286, 0.315517, 0.0338439, 0.0393744, 0.0339315, 0.154046]).reshape(1, 5, 6)
WARP_CTC_2_GRAD = np.array([
    -0.69824, 0.28562, 0.0831517, 0.0862751, 0.0816851, 0.161508,
    0.24082, -0.602467, 0.0557226, 0.0546814, 0.0557528, 0.19549,
    0.230246, 0.450868, 0.0389607, 0.038309, 0.0391602, -0.797544,
    0.280884, -0.570478, 0.0326593, 0.0339046, 0.0326856, 0.190345,
    -0.576714, 0.315517, 0.0338439, 0.0393744, 0.0339315, 0.154046]).reshape(1, 5, 6)
ASG_EMISSIONS = np.array([
    [[-0.4340, -0.0254, 0.3667, 0.4180, -0.3805, -0.1707],
     [0.1060, 0.3631, -0.1122, -0.3825, -0.0031, -0.3801],
     [0.0443, -0.3795, 0.3194, -0.3130, 0.0094, 0.1560],
     [0.1252, 0.2877, 0.1997, -0.4554, 0.2774, -0.2526],
     [-0.4001, -0.2402, 0.1295, 0.0172, 0.1805, -0.3299]],
    [[0.3298, -0.2259, -0.0959, 0.4909, 0.2996, -0.2543],
     [-0.2863, 0.3239, -0.3988, 0.0732, -0.2107, -0.4739],
     [-0.0906, 0.0480, -0.1301, 0.3975, -0.3317, -0.1967],
     [0.4372, -0.2006, 0.0094, 0.3281, 0.1873, -0.2945],
     [0.2399, 0.0320, -0.3768, -0.2849, -0.2248, 0.3186]],
    [[0.0225, -0.3867, -0.1929, -0.2904, -0.4958, -0.2533],
     [0.4001, -0.1517, -0.2799, -0.2915, 0.4198, 0.4506],
     [0.1446, -0.4753, -0.0711, 0.2876, -0.1851, -0.1066],
     [0.2081, -0.1190, -0.3902, -0.1668, 0.1911, -0.2848],
     [-0.3846, 0.1175, 0.1052, 0.2172, -0.0362, 0.3055]]])
ASG_LABELS = [[2, 1, 5, 1, 3], [4, 3, 5], [3, 2, 2, 1]]
ASG_GRAD = np.array([
    [[0.1060, 0.1595, -0.7639, 0.2485, 0.1118, 0.1380],
     [0.1915, -0.7524, 0.1539, 0.1175, 0.1717, 0.1178],
     [0.1738, 0.1137, 0.2288, 0.1216, 0.1678, -0.8057],
     [0.1766, -0.7923, 0.1902, 0.0988, 0.2056, 0.1210],
     [0.1212, 0.1422, 0.2059, -0.8160, 0.2166, 0.1300]],
    [[0.2029, 0.1164, 0.1325, 0.2383, -0.8032, 0.1131],
     [0.1414, 0.2602, 0.1263, -0.3441, -0.3009, 0.1172],
     [0.1557, 0.1788, 0.1496, -0.5498, 0.0140, 0.0516],
     [0.2306, 0.1219, 0.1503, -0.4244, 0.1796, -0.2579],
     [0.2149, 0.1745, 0.1160, 0.1271, 0.1350, -0.7675]],
    [[0.2195, 0.1458, 0.1770, -0.8395, 0.1307, 0.1666],
     [0.2148, 0.1237, -0.6613, -0.1223, 0.2191, 0.2259],
     [0.2002, 0.1077, -0.8386, 0.2310, 0.1440, 0.1557],
     [0.2197, -0.1466, -0.5742, 0.1510, 0.2160, 0.1342],
     [0.1050, -0.8265, 0.1714, 0.1917, 0.1488, 0.2094]]]) / 3
ASG_TRANS_GRAD = np.array([
    [0.3990, 0.3396, 0.3486, 0.3922, 0.3504, 0.3155],
    [0.3666, 0.0116, -1.6678, 0.3737, 0.3361, -0.7152],
    [0.3468, 0.3163, -1.1583, -0.6803, 0.3216, 0.2722],
    [0.3694, -0.6688, 0.3047, -0.8531, -0.6571, 0.2870],
    [0.3866, 0.3321, 0.3447, 0.3664, -0.2163, 0.3039],
    [0.3640, -0.6943, 0.2988, -0.6722, 0.3215, -0.1860]]) / 3

BACKENDS = ["gtn32", "gtn64"]


@pytest.fixture(params=BACKENDS)
def gtn(request):
    return request.getfixturevalue(request.param)


def test_ctc_trivial_and_uniform(gtn):
    # gtn_ctc_test.py:24-46
    with np.errstate(divide="ignore"):
        lp = np.log(np.array([1.0, 0.0, 0.0, 1.0, 1.0, 0.0]).reshape(1, 3, 2))
    assert abs(rc.ctc(gtn, lp, [[0, 0]], 1)["loss"]) < 1e-7
    lp = G.log_softmax(np.zeros((1, 3, 4)))
    assert abs(rc.ctc(gtn, lp, [[1, 2]], 3)["loss"] + math.log(0.25 ** 3 * 5)) < 1e-6


@pytest.mark.parametrize("probs,labels,loss,grad", [
    (WARP_CTC_1, [[0, 1, 2, 1, 0]], 3.34211, WARP_CTC_1_GRAD),
    (WARP_CTC_2, [[0, 1, 1, 0]], 5.42262, WARP_CTC_2_GRAD)])
def test_ctc_warpctc_vectors(gtn, probs, labels, loss, grad):
    # gtn_ctc_test.py:48-80, transducer_test.py:143-216: loss to 4 places and
    # the gradient w.r.t. the pre-softmax log-emissions (allclose defaults)
    logits = np.log(probs)
    res = rc.ctc(gtn, G.log_softmax(logits), labels, 5)
    assert abs(res["loss"] - loss) < 5e-5
    np.testing.assert_allclose(G.through_log_softmax(logits, res["grad"]), grad, rtol=1e-5, atol=1e-6)
    mine = dp_numpy.ctc(G.log_softmax(logits), labels, 5)
    assert abs(mine["loss"] - loss) < 5e-5
    np.testing.assert_allclose(mine["grad"], res["grad"], rtol=1e-4, atol=1e-6)


def test_asg_known_answer(gtn):
    # gtn_asg_test.py:25-105
    res = rc.asg(gtn, ASG_EMISSIONS, np.zeros((7, 6)), ASG_LABELS)
    assert abs(res["loss"] - 7.47995) < 5e-5
    np.testing.assert_allclose(res["grad"], ASG_GRAD, rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(res["grad_transitions"][1:], ASG_TRANS_GRAD, rtol=1e-3, atol=1e-5)
    mine = dp_numpy.asg(ASG_EMISSIONS, np.zeros((7, 6)), ASG_LABELS)
    assert abs(mine["loss"] - 7.47995) < 5e-5
    np.testing.assert_allclose(mine["grad"], res["grad"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(mine["grad_transitions"], res["grad_transitions"], rtol=1e-4, atol=1e-6)


def test_asg_viterbi_known_answer(gtn):
    # gtn_asg_test.py:107-124: raw best path [2, 1, 1, 0]
    em = np.array([0, 0, 0, 7, 0, 5, 4, 3, 0, 5, 8, 5, 0, 5, 4, 3], dtype=np.float32).reshape(1, 4, 4)
    tr = np.array([0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 0, 0, 0, 0, 2, 0, 2, 0, 0], dtype=np.float32).reshape(5, 4)
    path = rc.asg_viterbi(gtn, em, tr)[0]
    assert [p for i, p in enumerate(path) if i == 0 or p != path[i - 1]] == [3, 2, 1]


def test_stc_known_answers(gtn):
    # gtn_stc_test.py:25-51 through the module-level preprocessing (stc.py:196-220)
    def build(lp, targets):
        lp = np.transpose(lp, (1, 0, 2))
        lse = np.log(np.exp(lp[:, :, 1:]).sum(2, keepdims=True))
        sel = [0] + sorted(set(t for tg in targets for t in tg))
        tmap = {t: i for i, t in enumerate(sel)}
        lp = lp[:, :, sel]
        with np.errstate(divide="ignore", invalid="ignore"):
            neg = lse + np.log1p(1e-7 - np.exp(lp[:, :, 1:] - lse))
        return np.concatenate([lp, lse, neg], 2), [[tmap[t] for t in tg] for tg in targets]

    with np.errstate(divide="ignore"):
        lp = np.log(np.array([0.0, 1.0, 1.0, 0.0, 0.0, 1.0]).reshape(3, 1, 2))
    e, tg = build(lp, [[1, 1]])
    assert abs(rc.stc(gtn, e, tg, 1.0)["loss"]) < 1e-6
    lp = G.log_softmax(np.zeros((3, 1, 4)))
    e, tg = build(lp, [[1, 2]])
    assert abs(rc.stc(gtn, e, tg, 1.0)["loss"] + math.log(0.25 * 0.25 * 2.5)) < 1e-5


@pytest.mark.parametrize("case", ["small_none", "small_mean", "raw_mean", "repeats", "cfg1_raw", "cfg1_lsm"])
def test_ctc_fixtures(gtn, case):
    z = G.load("ctc")
    tg = G.unpack(z[case + "_targets"], z[case + "_offsets"])
    res = rc.ctc(gtn, z[case + "_emissions"], tg, int(z[case + "_blank"]), str(z[case + "_reduction"]))
    assert abs(res["loss"] - float(z[case + "_loss"])) <= 2e-5 * max(1.0, abs(float(z[case + "_loss"])))
    np.testing.assert_allclose(res["grad"], z[case + "_grad"], rtol=2e-3, atol=1e-6)
    if not case.startswith("cfg1"):
        mine = dp_numpy.ctc(z[case + "_emissions"], tg, int(z[case + "_blank"]), str(z[case + "_reduction"]))
        np.testing.assert_allclose(mine["grad"], z[case + "_grad"], rtol=2e-3, atol=1e-6)


def test_ctc_graph_indices_bit_exact(gtn):
    z = G.load("ctc")
    g = rc.ctc_graph(gtn, [3, 3, 1, 0, 0, 2], 5)
    mine, ref = rc.graph_arrays(g), G.graph_of(z, "graph")
    for k in ref:
        np.testing.assert_array_equal(mine[k], ref[k])
    np.testing.assert_array_equal(np.array(g.in_order()), z["graph_in_order"])
    np.testing.assert_array_equal(np.array(g.out_order()), z["graph_out_order"])


@pytest.mark.parametrize("case", ["small_none", "small_mean", "mid_mean"])
def test_asg_fixtures(gtn, case):
    z = G.load("asg")
    tg = G.unpack(z[case + "_targets"], z[case + "_offsets"])
    res = rc.asg(gtn, z[case + "_emissions"], z[case + "_transitions"], tg, str(z[case + "_reduction"]))
    assert abs(res["loss"] - float(z[case + "_loss"])) <= 2e-5 * max(1.0, abs(float(z[case + "_loss"])))
    np.testing.assert_allclose(res["grad"], z[case + "_grad"], rtol=2e-3, atol=2e-6)
    np.testing.assert_allclose(res["grad_transitions"], z[case + "_grad_transitions"], rtol=2e-3, atol=2e-6)
    mine = dp_numpy.asg(z[case + "_emissions"], z[case + "_transitions"], tg, str(z[case + "_reduction"]))
    np.testing.assert_allclose(mine["grad"], z[case + "_grad"], rtol=2e-3, atol=2e-6)
    np.testing.assert_allclose(mine["grad_transitions"], z[case + "_grad_transitions"], rtol=2e-3, atol=2e-6)


def test_asg_graph_indices_bit_exact(gtn):
    z = G.load("asg")
    g = rc.asg_force_align_graph(gtn, [2, 0, 0, 1])
    mine, ref = rc.graph_arrays(g), G.graph_of(z, "falgraph")
    for k in ref:
        np.testing.assert_array_equal(mine[k], ref[k])
    g = rc.asg_transitions_graph(gtn, np.zeros((4, 3), dtype=np.float32))
    mine, ref = rc.graph_arrays(g), G.graph_of(z, "transgraph")
    for k in ("start", "accept", "src", "dst", "ilabel", "olabel"):
        np.testing.assert_array_equal(mine[k], ref[k])


@pytest.mark.parametrize("case", ["fn_none", "fn_mean"])
def test_stc_fixtures(gtn, case):
    z = G.load("stc")
    tg = G.unpack(z[case + "_targets"], z[case + "_offsets"])
    res = rc.stc(gtn, z[case + "_emissions"], tg, float(z[case + "_prob"]), str(z[case + "_reduction"]))
    assert abs(res["loss"] - float(z[case + "_loss"])) <= 2e-5 * max(1.0, abs(float(z[case + "_loss"])))
    np.testing.assert_allclose(res["grad"], z[case + "_grad"], rtol=2e-3, atol=2e-6)


def test_stc_graph_indices_bit_exact(gtn):
    z = G.load("stc")
    mine, ref = rc.graph_arrays(rc.stc_graph(gtn, [2, 1, 1], 4, 0.5)), G.graph_of(z, "graph")
    for k in ("start", "accept", "src", "dst", "ilabel", "olabel"):
        np.testing.assert_array_equal(mine[k], ref[k])
    np.testing.assert_allclose(mine["weight"], ref["weight"], rtol=1e-6)


@pytest.mark.parametrize("name,blank,rep", [("wp_none", "none", True), ("wp_opt", "optional", True),
                                            ("wp_norep", "optional", False), ("wp_forced", "forced", True)])
def test_transducer_wordpiece_fixtures(gtn, name, blank, rep):
    z = G.load("transducer")
    tokens, g2i = ["a", "b", "ab", "ba", "aba"], {"a": 0, "b": 1}
    tg = G.unpack(z["wp_targets"], z["wp_offsets"])
    crit = rc.Transducer(gtn, tokens, g2i, blank=blank, allow_repeats=rep, reduction="mean")
    logits = z[name + "_logits"]
    res = crit.loss(G.log_softmax(logits), tg)
    assert abs(res["loss"] - float(z[name + "_loss"])) <= 2e-5 * max(1.0, abs(float(z[name + "_loss"])))
    np.testing.assert_allclose(G.through_log_softmax(logits, res["grad"]), z[name + "_grad_logits"],
                               rtol=2e-3, atol=2e-6)
    assert crit.viterbi(logits) == G.unpack(z[name + "_viterbi"], z[name + "_viterbi_offsets"])
    crit.tokens.arc_sort(True)
    for prefix, g in (("_tokens", crit.tokens), ("_lexicon", crit.lexicon),
                      ("_align1", crit.alignment_graph(tg[1]))):
        mine, ref = rc.graph_arrays(g), G.graph_of(z, name + prefix)
        for k in ("start", "accept", "src", "dst", "ilabel", "olabel"):
            np.testing.assert_array_equal(mine[k], ref[k])


@pytest.mark.parametrize("name,ngram,blank,rep", [("ngram1", 1, "optional", False),
                                                  ("ngram2", 2, "optional", False),
                                                  ("ngram2_asg", 2, "none", True)])
def test_transducer_ngram_fixtures(gtn, name, ngram, blank, rep):
    z = G.load("transducer")
    N = 4
    crit = rc.Transducer(gtn, [(i,) for i in range(N)], {i: i for i in range(N)}, ngram=ngram,
                         blank=blank, allow_repeats=rep, reduction="mean")
    tg = G.unpack(z[name + "_targets"], z[name + "_offsets"])
    res = crit.loss(z[name + "_emissions"], tg, z[name + "_params"])
    assert abs(res["loss"] - float(z[name + "_loss"])) <= 2e-5 * max(1.0, abs(float(z[name + "_loss"])))
    np.testing.assert_allclose(res["grad"], z[name + "_grad"], rtol=2e-3, atol=2e-6)
    np.testing.assert_allclose(res["grad_transitions"], z[name + "_grad_params"], rtol=2e-3, atol=2e-6)
    assert crit.viterbi(z[name + "_emissions"], z[name + "_params"]) == \
        G.unpack(z[name + "_viterbi"], z[name + "_viterbi_offsets"])
    mine, ref = rc.graph_arrays(crit.transitions), G.graph_of(z, name + "_transitions")
    for k in ("start", "accept", "src", "dst", "ilabel", "olabel"):
        np.testing.assert_array_equal(mine[k], ref[k])


def _graph_from_arrays(gtn, a):
    g = gtn.Graph(True)
    for s, acc in zip(a["start"], a["accept"]):
        g.add_node(bool(s), bool(acc))
    for s, d, i, o, w in zip(a["src"], a["dst"], a["ilabel"], a["olabel"], a["weight"]):
        g.add_arc(int(s), int(d), int(i), int(o), float(w))
    return g


def test_transducer_backoff_fixture(gtn):
    z = G.load("transducer")
    N = 5
    trans = _graph_from_arrays(gtn, G.graph_of(z, "backoff_file"))
    crit = rc.Transducer(gtn, [(i,) for i in range(N)], {i: i for i in range(N)}, blank="optional",
                         allow_repeats=False, transitions=trans)
    tg = G.unpack(z["backoff_targets"], z["backoff_offsets"])
    res = crit.loss(z["backoff_emissions"], tg, z["backoff_params"])
    assert abs(res["loss"] - float(z["backoff_loss"])) <= 2e-5 * max(1.0, abs(float(z["backoff_loss"])))
    np.testing.assert_allclose(res["grad"], z["backoff_grad"], rtol=2e-3, atol=2e-6)
    np.testing.assert_allclose(res["grad_transitions"], z["backoff_grad_params"], rtol=2e-3, atol=2e-6)
