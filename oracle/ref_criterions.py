"""ORACLE — TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's criterion Functions on top of the ``gtn``
shim (``oracle/gtn`` float32 — GTN's scalar type — or ``oracle/gtn64`` float64,
the truth build).  Each function follows the reference flow it cites: build
the per-utterance graphs, compose with the emissions graph, log-semiring
forward score, reverse sweep.  They return plain numbers / numpy arrays: the
batch-mean loss and the gradients ``backward`` would hand to autograd for
``grad_output = 1``.

``/root/reference`` does not exist on the GPU box, so the GPU parity tests use
THESE functions; tests/test_oracle_reference.py checks, in the build container,
that they agree with the reference's own ``criterions/*.py`` (imported
unchanged against the same shim) and with its golden values.

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs
import this module.
"""
import itertools
import math

import numpy as np


def _f32(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float32))


def _emissions_graph(gtn, e_b, calc_grad=True):
    """linear_graph(T, C) + set_weights (ctc.py:40-44; asg.py:96-100;
    stc.py:74-78; transducer.py:262-264)."""
    T, C = e_b.shape
    g = gtn.linear_graph(T, C, gtn.Device(gtn.CPU), calc_grad)
    buf = _f32(e_b)
    g.set_weights(buf.ctypes.data)
    return g


def _scale(reduction, n):
    if reduction == "mean":
        return 1.0 / n if n > 0 else 1.0
    if reduction != "none":
        raise ValueError("invalid value for reduction '" + str(reduction) + "'")
    return 1.0


# --------------------------------------------------------------------------
# CTC  (criterions/ctc.py)
# --------------------------------------------------------------------------
def ctc_graph(gtn, target, blank):
    """ctc.py:15-29 — S = 2L+1 nodes; per node l, in this order: self loop,
    arc from l-1, skip arc from l-2 when l is odd and the label differs from
    the previous label; then arc_sort on ilabel."""
    g = gtn.Graph(False)
    n_states = 2 * len(target) + 1
    for s in range(n_states):
        k = (s - 1) // 2
        g.add_node(s == 0, s >= n_states - 2)
        lab = target[k] if s % 2 else blank
        g.add_arc(s, s, lab)
        if s > 0:
            g.add_arc(s - 1, s, lab)
        if s % 2 and s > 1 and lab != target[k - 1]:
            g.add_arc(s - 2, s, lab)
    g.arc_sort(False)
    return g


def ctc(gtn, emissions, targets, blank=0, reduction="none", want_grad=True):
    """CTCLossFunction.forward/backward (ctc.py:31-94)."""
    emissions = _f32(emissions)
    B, T, C = emissions.shape
    losses = [None] * B
    scales = [None] * B
    graphs = [None] * B

    def fwd(b):
        g_em = _emissions_graph(gtn, emissions[b], want_grad)
        g_crit = ctc_graph(gtn, targets[b], blank)
        losses[b] = gtn.negate(gtn.forward_score(gtn.intersect(g_em, g_crit)))
        scales[b] = _scale(reduction, len(targets[b]))
        graphs[b] = g_em

    gtn.parallel_for(fwd, range(B))
    per_utt = np.array([losses[b].item() * scales[b] for b in range(B)], dtype=np.float64)
    out = {"loss": float(per_utt.mean()), "losses": per_utt}
    if want_grad:
        grad = np.zeros((B, T, C), dtype=np.float64)

        def bwd(b):
            gtn.backward(losses[b], False)
            if graphs[b].has_grad():
                grad[b] = graphs[b].grad().weights_to_numpy().reshape(T, C) * scales[b]

        gtn.parallel_for(bwd, range(B))
        out["grad"] = grad / B
    return out


# --------------------------------------------------------------------------
# ASG  (criterions/asg.py)
# --------------------------------------------------------------------------
def asg_transitions_graph(gtn, transitions, calc_grad=False):
    """asg.py:53-69 — node 0 start; nodes 1..C accept; arc i-1 = (0 -> i,
    label i-1); arc C + i*C + j = (j+1 -> i+1, label i); weights are
    transitions.flatten(), i.e. transitions[0, i] = score(i | <s>) and
    transitions[1+i, j] = score(i | prev = j)."""
    transitions = _f32(transitions)
    C = transitions.shape[1]
    assert transitions.shape == (C + 1, C)
    g = gtn.Graph(calc_grad)
    g.add_node(True)
    for i in range(1, C + 1):
        g.add_node(False, True)
        g.add_arc(0, i, i - 1)
    for i in range(C):
        for j in range(C):
            g.add_arc(j + 1, i + 1, i)
    g.set_weights(transitions.ctypes.data)
    g.mark_arc_sorted(False)
    g.mark_arc_sorted(True)
    return g


def asg_force_align_graph(gtn, target):
    """asg.py:71-81 — chain 0..L with a self loop on every non-start node."""
    g = gtn.Graph(False)
    L = len(target)
    g.add_node(True)
    for k in range(1, L + 1):
        g.add_node(False, k == L)
        g.add_arc(k - 1, k, target[k - 1])
        g.add_arc(k, k, target[k - 1])
    g.arc_sort(True)
    return g


def asg(gtn, emissions, transitions, targets, reduction="none", want_grad=True):
    """ASGLossFunction.forward/backward (asg.py:83-185): loss_b =
    Z(em ∘ trans) − Z((fal ∘ trans) ∘ em)."""
    emissions = _f32(emissions)
    transitions = _f32(transitions)
    B, T, C = emissions.shape
    losses = [None] * B
    scales = [None] * B
    ems = [None] * B
    trs = [None] * B

    def fwd(b):
        g_em = _emissions_graph(gtn, emissions[b], want_grad)
        g_tr = asg_transitions_graph(gtn, transitions, want_grad)
        g_fal = asg_force_align_graph(gtn, targets[b])
        fal = gtn.forward_score(gtn.intersect(gtn.intersect(g_fal, g_tr), g_em))
        fcc = gtn.forward_score(gtn.intersect(g_em, g_tr))
        losses[b] = gtn.subtract(fcc, fal)
        scales[b] = _scale(reduction, len(targets[b]))
        ems[b] = g_em
        trs[b] = g_tr

    gtn.parallel_for(fwd, range(B))
    per_utt = np.array([losses[b].item() * scales[b] for b in range(B)], dtype=np.float64)
    out = {"loss": float(per_utt.mean()), "losses": per_utt}
    if want_grad:
        g_em_all = np.zeros((B, T, C), dtype=np.float64)
        g_tr_all = np.zeros((B, C + 1, C), dtype=np.float64)

        def bwd(b):
            gtn.backward(losses[b], False)
            if ems[b].has_grad():
                g_em_all[b] = ems[b].grad().weights_to_numpy().reshape(T, C) * scales[b]
            if trs[b].has_grad():
                g_tr_all[b] = trs[b].grad().weights_to_numpy().reshape(C + 1, C) * scales[b]

        gtn.parallel_for(bwd, range(B))
        out["grad"] = g_em_all / B
        out["grad_transitions"] = g_tr_all.mean(0)
    return out


def asg_viterbi(gtn, emissions, transitions):
    """ASG.viterbi core (asg.py:217-226): best label path through em ∘ trans,
    before collapsing / replabel unpacking."""
    emissions = _f32(emissions)
    B = emissions.shape[0]
    paths = [None] * B

    def run(b):
        g_em = _emissions_graph(gtn, emissions[b], False)
        g_tr = asg_transitions_graph(gtn, transitions)
        paths[b] = gtn.viterbi_path(gtn.intersect(g_em, g_tr)).labels_to_list()

    gtn.parallel_for(run, range(B))
    return paths


# --------------------------------------------------------------------------
# STC  (criterions/stc.py)
# --------------------------------------------------------------------------
def stc_graph(gtn, target, star_idx, prob):
    """stc.py:22-64 — CTC-like chain with self loops only on blank states and
    unconditional skip arcs, plus one <star> node per gap carrying log(prob)
    on its entering / looping arcs."""
    g = gtn.Graph(False)
    L = len(target)
    n_states = 2 * L + 1
    for s in range(n_states):
        k = (s - 1) // 2
        g.add_node(s == 0, s >= n_states - 2)
        lab = target[k] if s % 2 else 0
        if lab == 0:
            g.add_arc(s, s, lab)
        if s > 0:
            g.add_arc(s - 1, s, lab)
        if s % 2 and s > 1:
            g.add_arc(s - 2, s, lab)
    lp = math.log(prob)
    for k in range(L + 1):
        prev_tok, prev_blank = 2 * k - 1, 2 * k
        c = g.add_node(False, k == L)
        idx = star_idx if k == L else star_idx + target[k]
        if prev_tok >= 0:
            g.add_arc(prev_tok, c, idx, idx, lp)
        g.add_arc(prev_blank, c, idx, idx, lp)
        g.add_arc(c, c, idx, idx, lp)
        if k < L:
            g.add_arc(c, 2 * k + 1, target[k])
        g.add_arc(c, prev_blank, 0)
    return g


def stc(gtn, emissions, targets, prob, reduction="none", want_grad=True):
    """STCLossFunction.forward/backward (stc.py:66-129); criterion graph is
    the FIRST compose operand; "mean" divides by T."""
    emissions = _f32(emissions)
    B, T, Cstar = emissions.shape
    star = Cstar // 2
    losses = [None] * B
    graphs = [None] * B
    scale = _scale(reduction, T)

    def fwd(b):
        g_em = _emissions_graph(gtn, emissions[b], want_grad)
        g_crit = stc_graph(gtn, targets[b], star, prob)
        g_crit.arc_sort(False)
        losses[b] = gtn.negate(gtn.forward_score(gtn.compose(g_crit, g_em)))
        graphs[b] = g_em

    gtn.parallel_for(fwd, range(B))
    per_utt = np.array([losses[b].item() * scale for b in range(B)], dtype=np.float64)
    out = {"loss": float(per_utt.mean()), "losses": per_utt}
    if want_grad:
        grad = np.zeros((B, T, Cstar), dtype=np.float64)

        def bwd(b):
            gtn.backward(losses[b], False)
            if graphs[b].has_grad():
                grad[b] = graphs[b].grad().weights_to_numpy().reshape(T, Cstar) * scale

        gtn.parallel_for(bwd, range(B))
        out["grad"] = grad / B
    return out


# --------------------------------------------------------------------------
# Transducer  (criterions/transducer.py)
# --------------------------------------------------------------------------
def chain_graph(gtn, seq):
    """transducer.py:23-29."""
    g = gtn.Graph(False)
    g.add_node(True)
    for i, s in enumerate(seq):
        g.add_node(False, i == len(seq) - 1)
        g.add_arc(i, i + 1, int(s))
    return g


def ngram_transitions_graph(gtn, ngram, num_tokens, calc_grad=False):
    """transducer.py:32-58 — dense n-gram acceptor; for ngram > 1 a final
    </s> node reached by epsilon arcs from every other node."""
    g = gtn.Graph(calc_grad)
    g.add_node(True, ngram == 1)
    state = {(): 0}
    for n in range(1, ngram):
        for hist in itertools.product(range(num_tokens), repeat=n):
            src = state[hist[:-1]]
            dst = g.add_node(False, ngram == 1)
            state[hist] = dst
            g.add_arc(src, dst, hist[-1])
    for hist in itertools.product(range(num_tokens), repeat=ngram):
        g.add_arc(state[hist[:-1]], state[hist[1:]], hist[-1])
    if ngram > 1:
        end = g.add_node(False, True)
        for src in range(end):
            g.add_arc(src, end, gtn.epsilon)
    return g


def lexicon_graph(gtn, word_pieces, graphemes_to_idx):
    """transducer.py:61-75 — letters -> word-piece transducer, one loop
    through node 0 per word piece, output label on the last letter."""
    g = gtn.Graph(False)
    g.add_node(True, True)
    for i, wp in enumerate(word_pieces):
        prev = 0
        for ch in wp[:-1]:
            n = g.add_node()
            g.add_arc(prev, n, graphemes_to_idx[ch], gtn.epsilon)
            prev = n
        g.add_arc(prev, 0, graphemes_to_idx[wp[-1]], i)
    g.arc_sort()
    return g


def token_graph(gtn, token_list, blank="none", allow_repeats=True):
    """transducer.py:78-123 — per-token emission models (self loops) with
    none / optional / forced blank and optional no-repeat constraint."""
    if not allow_repeats and blank != "optional":
        raise ValueError("Must use blank='optional' if disallowing repeats.")
    n = len(token_list)
    g = gtn.Graph(False)
    g.add_node(True, True)
    for _ in range(n):
        g.add_node(False, blank != "forced")
    if blank != "none":
        g.add_node()
        g.add_arc(0, n + 1, n, gtn.epsilon)
        g.add_arc(n + 1, 0, gtn.epsilon)
    for i in range(n):
        g.add_arc((n + 1) if blank == "forced" else 0, i + 1, i)
        g.add_arc(i + 1, i + 1, i, gtn.epsilon)
        if allow_repeats:
            if blank == "forced":
                g.add_arc(i + 1, n + 1, n, gtn.epsilon)
            else:
                g.add_arc(i + 1, 0, gtn.epsilon)
        else:
            g.add_arc(i + 1, n + 1, n, gtn.epsilon)
            for j in range(n):
                if i != j:
                    g.add_arc(i + 1, j + 1, j, j)
    return g


class Transducer:
    """Transducer module + TransducerLossFunction (transducer.py:126-348)."""

    def __init__(self, gtn, tokens, graphemes_to_idx, ngram=0, transitions=None,
                 blank="none", allow_repeats=True, reduction="none"):
        if blank not in ("optional", "forced", "none"):
            raise ValueError("Invalid value specificed for blank.")
        self.gtn = gtn
        self.tokens = token_graph(gtn, tokens, blank, allow_repeats)
        self.lexicon = lexicon_graph(gtn, tokens, graphemes_to_idx)
        if ngram > 0 and transitions is not None:
            raise ValueError("Only one of ngram and transitions may be specified")
        if ngram > 0:
            transitions = ngram_transitions_graph(
                gtn, ngram, len(tokens) + int(blank != "none"), True)
        self.transitions = transitions
        if transitions is not None:
            self.transitions.arc_sort()
        self.reduction = reduction

    def alignment_graph(self, target):
        """transducer.py:265-276 (+279-281 with transitions)."""
        gtn = self.gtn
        tgt = chain_graph(gtn, target)
        tgt.arc_sort(True)
        tokens_target = gtn.remove(gtn.project_output(gtn.compose(tgt, self.lexicon)))
        tokens_target.arc_sort()
        align = gtn.project_input(gtn.remove(gtn.compose(self.tokens, tokens_target)))
        align.arc_sort()
        return align

    def loss(self, emissions, targets, transition_params=None, want_grad=True,
             log_softmax_when_unnormalised=False):
        gtn = self.gtn
        emissions = _f32(emissions)
        B, T, C = emissions.shape
        trans = self.transitions
        if trans is not None:
            if transition_params is None:
                raise ValueError("Specified transitions, but not transition params.")
            tp = _f32(transition_params)
            trans.set_weights(tp.ctypes.data)
            trans.calc_grad = want_grad
            trans.zero_grad()
        self.tokens.arc_sort(True)
        losses = [None] * B
        graphs = [None] * B

        def fwd(b):
            g_em = _emissions_graph(gtn, emissions[b], want_grad)
            align = self.alignment_graph(targets[b])
            if trans is not None:
                align = gtn.intersect(trans, align)
                align.arc_sort()
            score = gtn.forward_score(gtn.intersect(g_em, align))
            if trans is not None:
                score = gtn.subtract(score, gtn.forward_score(gtn.intersect(g_em, trans)))
            losses[b] = gtn.negate(score)
            graphs[b] = g_em

        gtn.parallel_for(fwd, range(B))
        scales = [_scale(self.reduction, len(t)) for t in targets]
        per_utt = np.array([losses[b].item() * scales[b] for b in range(B)], dtype=np.float64)
        out = {"loss": float(per_utt.mean()), "losses": per_utt}
        if want_grad:
            grad = np.zeros((B, T, C), dtype=np.float64)

            def bwd(b):
                gtn.backward(losses[b], gtn.scalar_graph(scales[b]))
                if graphs[b].has_grad():
                    grad[b] = graphs[b].grad().weights_to_numpy().reshape(T, C)

            gtn.parallel_for(bwd, range(B))
            out["grad"] = grad / B
            if trans is not None and trans.has_grad():
                out["grad_transitions"] = trans.grad().weights_to_numpy().astype(np.float64) / B
        return out

    def viterbi(self, emissions, transition_params=None):
        """Transducer.viterbi (transducer.py:199-234)."""
        gtn = self.gtn
        emissions = _f32(emissions)
        B = emissions.shape[0]
        trans = self.transitions
        if trans is not None:
            tp = _f32(transition_params)
            trans.set_weights(tp.ctypes.data)
            trans.calc_grad = False
        self.tokens.arc_sort()
        paths = [None] * B

        def run(b):
            g_em = _emissions_graph(gtn, emissions[b], False)
            full = gtn.intersect(g_em, trans) if trans is not None else g_em
            path = gtn.remove(gtn.viterbi_path(full))
            path = gtn.compose(path, self.tokens)
            path = gtn.viterbi_path(path)
            path = gtn.remove(gtn.project_output(path))
            paths[b] = path.labels_to_list()

        gtn.parallel_for(run, range(B))
        return paths


# --------------------------------------------------------------------------
# structure dumps (index-exact comparisons)
# --------------------------------------------------------------------------
def graph_arrays(g):
    """(start flags, accept flags, src, dst, ilabel, olabel, weights) of a
    shim graph, as numpy arrays — what "arc/state indices bit-exact" is
    checked on."""
    n = g.num_nodes()
    start = np.zeros(n, dtype=np.int32)
    accept = np.zeros(n, dtype=np.int32)
    start[g.start()] = 1
    accept[g.accept()] = 1
    src, dst, il, ol = g.arcs_to_numpy()
    return {
        "start": start, "accept": accept,
        "src": np.asarray(src, dtype=np.int32), "dst": np.asarray(dst, dtype=np.int32),
        "ilabel": np.asarray(il, dtype=np.int32), "olabel": np.asarray(ol, dtype=np.int32),
        "weight": np.asarray(g.weights_to_numpy(), dtype=np.float64),
    }
