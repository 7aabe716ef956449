"""Generic WFST transducer criterion on B200 — same interface as the reference's
criterions/transducer.py (criterion part).

``TransducerLossFunction.forward(ctx, inputs, targets, tokens, lexicon,
transition_params=None, transitions=None, reduction="none")`` / ``.backward`` ->
7-tuple with gradients at positions 0 and 4 mirror transducer.py:239-248,340-348;
``Transducer(tokens, graphemes_to_idx, ngram=0, transitions=None, blank="none",
allow_repeats=True, reduction="none")`` keeps the ``transition_params`` parameter
(transducer.py:149-183; part of the checkpoint format, train.py:117).

Per utterance the reference builds, in Python under the GIL, target o lexicon, the token
decomposition DAG and the alignment acceptor (transducer.py:265-276), intersects it with
the emissions graph and runs forward_score / backward on the CPU.  Here the host graphs
of the whole batch are built in C++ on a thread pool (wfst_transducer_alignment_graphs),
packed once, and every utterance's lattice is scored by one launch of the generic
lattice kernel; with a transition graph the normaliser Z(emissions o transitions) is a
second launch over a single shared graph."""
import itertools

import collections

import numpy as np
import torch

from .. import _lib, _runtime as rt
from .. import graph as G
from ..lattice import lattice_forward_backward


def make_scalar_graph(weight):
    """transducer.py:15-20"""
    g = G.Graph()
    g.add_node(True)
    g.add_node(False, True)
    g.add_arc(0, 1, 0, 0, weight)
    return g


def make_chain_graph(sequence):
    """transducer.py:23-29"""
    g = G.Graph(False)
    g.add_node(True)
    for i, s in enumerate(sequence):
        g.add_node(False, i == len(sequence) - 1)
        g.add_arc(i, i + 1, int(s))
    return g


def make_transitions_graph(ngram, num_tokens, calc_grad=False):
    """Dense n-gram acceptor (transducer.py:32-58); for ngram > 1 a final </s> node is
    reached by epsilon arcs from every other node."""
    g = G.Graph(calc_grad)
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
            g.add_arc(src, end, G.epsilon)
    return g


def make_lexicon_graph(word_pieces, graphemes_to_idx):
    """Letters -> word-piece transducer (transducer.py:61-75): one loop through node 0
    per word piece, the output label on its last letter."""
    g = G.Graph(False)
    g.add_node(True, True)
    for i, wp in enumerate(word_pieces):
        prev = 0
        for ch in wp[:-1]:
            n = g.add_node()
            g.add_arc(prev, n, graphemes_to_idx[ch], G.epsilon)
            prev = n
        g.add_arc(prev, 0, graphemes_to_idx[wp[-1]], i)
    g.arc_sort()
    return g


def make_token_graph(token_list, blank="none", allow_repeats=True):
    """Per-token emission models (transducer.py:78-123).  The O(V^2) no-repeat arcs are
    added in bulk."""
    if not allow_repeats and blank != "optional":
        raise ValueError("Must use blank='optional' if disallowing repeats.")
    n = len(token_list)
    g = G.Graph(False)
    g.add_node(True, True)
    for _ in range(n):
        g.add_node(False, blank != "forced")
    if blank != "none":
        g.add_node()
        g.add_arc(0, n + 1, n, G.epsilon)
        g.add_arc(n + 1, 0, G.epsilon)
    if allow_repeats:
        for i in range(n):
            g.add_arc((n + 1) if blank == "forced" else 0, i + 1, i)
            g.add_arc(i + 1, i + 1, i, G.epsilon)
            if blank == "forced":
                g.add_arc(i + 1, n + 1, n, G.epsilon)
            else:
                g.add_arc(i + 1, 0, G.epsilon)
        return g
    # no repeats: per token i, in the reference's arc order: (0 -> i+1, i), (i+1 -> i+1, i:eps),
    # (i+1 -> blank, n:eps), then (i+1 -> j+1, j) for every j != i
    per = 3 + (n - 1)
    src = np.empty(n * per, dtype=np.int32)
    dst = np.empty_like(src)
    il = np.empty_like(src)
    ol = np.empty_like(src)
    others = np.arange(n, dtype=np.int32)
    for i in range(n):
        o = i * per
        src[o:o + 3] = (0, i + 1, i + 1)
        dst[o:o + 3] = (i + 1, i + 1, n + 1)
        il[o:o + 3] = (i, i, n)
        ol[o:o + 3] = (i, G.epsilon, G.epsilon)
        js = others[others != i]
        src[o + 3:o + per] = i + 1
        dst[o + 3:o + per] = js + 1
        il[o + 3:o + per] = js
        ol[o + 3:o + per] = js
    g.add_arcs(src, dst, il, ol)
    return g


def _graph_has_epsilon(g):
    cached = getattr(g, "_has_eps", None)
    if cached is None:
        from ..epsilon import has_epsilon
        cached = has_epsilon(g.arrays())
        try:
            g._has_eps = cached
        except AttributeError:
            pass
    return cached


def _fold_batch(transitions, align_handles, B, dev):
    """(packed, ties) of intersect(transitions, alignment_b) for the whole batch, epsilons folded: the
    composition, the fold and the tie arrays come from the host library in one call on host threads
    (wfst_fold_transitions_batch); align_handles = None folds the transition graph itself."""
    import ctypes
    from ..epsilon import FoldedBatch
    L = _lib.lib()
    out = (ctypes.c_int32 * B)()
    fold = L.wfst_fold_transitions_batch(transitions._h, align_handles, B, out)
    if fold < 0:
        _lib.check(fold)
    try:
        ties = FoldedBatch.from_fold(fold, dev)
        packed = G.pack_handles(out, B, dev)
    finally:
        L.wfst_fold_destroy(fold)
        G.destroy_handles(out, B)
    return packed, ties


# (packed, ties) of intersect(transitions, alignments) for batches met again (the reference
# benchmark repeats its batch; fixed mini-batches over epochs): the structure depends on the graphs
# and the targets only — the weights are gathered from transition_params on every call
_FOLDED_LRU = collections.OrderedDict()


def _fold_batch_lru(transitions, align_handles, B, dev, batch_key):
    if batch_key is None:
        return _fold_batch(transitions, align_handles, B, dev)
    tokens, lexicon, flat, offs = batch_key
    key = (id(tokens), tokens._h, tokens.num_arcs(), id(lexicon), lexicon._h, lexicon.num_arcs(),
           id(transitions), transitions._h, transitions.num_arcs(), str(dev), B,
           hash(flat.tobytes()), hash(offs.tobytes()), int(flat.size))
    hit = _FOLDED_LRU.get(key)
    if hit is not None and np.array_equal(hit[2], flat) and np.array_equal(hit[3], offs):
        _FOLDED_LRU.move_to_end(key)
        return hit[0], hit[1]
    packed, ties = _fold_batch(transitions, align_handles, B, dev)
    _FOLDED_LRU[key] = (packed, ties, flat.copy(), offs.copy())
    while len(_FOLDED_LRU) > _PACKED_LRU_MAX:
        _FOLDED_LRU.popitem(last=False)
    return packed, ties


def _folded_shared(transitions, dev):
    """(packed, ties) of the epsilon-folded transition graph, cached on the Graph object
    (the topology does not change between steps; the weights are gathered per call)."""
    key = "_folded_%s" % str(dev)
    hit = getattr(transitions, key, None)
    if hit is None:
        hit = _fold_batch(transitions, None, 1, dev)
        try:
            setattr(transitions, key, hit)
        except AttributeError:
            pass
    return hit


def _forward_with_epsilon_transitions(e, align_handles, transitions, transition_params, sc, need_e, need_t,
                                      batch_key=None):
    """TransducerLossFunction.forward (transducer.py:279-309) when the transition graph has
    epsilon arcs (ngram > 1: the </s> arcs of make_transitions_graph :52-56; loaded back-off
    graphs).  intersect(transitions, alignments) is done by the host library with the
    epsilons in place, exactly as the reference does; then both the composed graphs and the
    transition graph itself are folded (epsilon.py) and scored by the lattice kernel with
    final weights.  Gradients return to `transition_params` through the fold's provenance."""
    B = e.shape[0]
    dev = e.device
    tp = transition_params.detach().to(dev, torch.float32).contiguous()
    packed, ties = _fold_batch_lru(transitions, align_handles, B, dev, batch_key)
    w, fw, pw = ties.weights(tp)
    gs = -sc / B
    z_align, g_e, g_w, g_f = lattice_forward_backward(
        e, packed, grad_scale=gs, want_grad_emissions=need_e, want_grad_weights=need_t,
        weights=w, final_weights=fw)
    spacked, sties = _folded_shared(transitions, dev)
    sw, sfw, spw = sties.weights(tp)
    z_norm, _, g_wn, g_fn = lattice_forward_backward(
        e, spacked, grad_scale=-gs, want_grad_emissions=False, want_grad_weights=need_t,
        weights=sw, shared=True, accumulate_into=g_e if need_e else None, final_weights=sfw)
    g_tp = None
    if need_t:
        g_tp = ties.scatter_grads(tp.numel(), g_w, g_f, fw, pw) + \
            sties.scatter_grads(tp.numel(), g_wn, g_fn, sfw, spw)
    loss = (-(z_align - z_norm) * sc).mean()
    return loss, g_e, g_tp


# Device-resident packed batches of alignment graphs, keyed by (token graph, lexicon graph, the
# batch's targets): a batch met again (the reference benchmark repeats its batch every iteration;
# fixed mini-batches over epochs) is neither packed nor copied to the device again.  A few
# entries, least recently used out first; the graphs themselves come from the host library's
# alignment cache (wfst_transducer_alignment_cache), so a hit here skips pack + H2D only.
_PACKED_LRU = collections.OrderedDict()
_PACKED_LRU_MAX = 8


def _packed_batch(tokens, lexicon, flat, offs, handles, B, dev):
    key = (id(tokens), tokens._h, tokens.num_arcs(), id(lexicon), lexicon._h, lexicon.num_arcs(), str(dev), B,
           hash(flat.tobytes()), hash(offs.tobytes()), int(flat.size))
    hit = _PACKED_LRU.get(key)
    if hit is not None and np.array_equal(hit[1], flat) and np.array_equal(hit[2], offs):
        _PACKED_LRU.move_to_end(key)
        G.destroy_handles(handles, B)
        return hit[0]
    packed = G.pack_handles(handles, B, dev)
    G.destroy_handles(handles, B)
    _PACKED_LRU[key] = (packed, flat.copy(), offs.copy())
    while len(_PACKED_LRU) > _PACKED_LRU_MAX:
        _PACKED_LRU.popitem(last=False)
    return packed


class TransducerLossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inputs, targets, tokens, lexicon, transition_params=None, transitions=None,
                reduction="none"):
        B, T, C = inputs.shape
        rt.require_cuda(inputs, "inputs")
        if transitions is not None and transition_params is None:
            raise ValueError("Specified transitions, but not transition params.")
        if reduction == "mean":
            scales = [(1.0 / len(t) if len(t) > 0 else 1.0) for t in targets]
        else:
            scales = [1.0] * B      # like the reference, anything but "mean" is "none" (transducer.py:302-305)
        e = rt.to_device(inputs.detach())
        dev = e.device
        L = _lib.lib()
        flat, offs = rt.flatten_targets_host(targets)
        import ctypes
        handles = (ctypes.c_int32 * B)()
        _lib.check(L.wfst_transducer_alignment_graphs(
            tokens._h, lexicon._h, flat.ctypes.data, offs.ctypes.data, B, handles))
        # without a transition graph the alignment graphs are only packed and freed: no Python
        # wrapper (and no per-graph destructor call) for them
        aligns = [G.Graph(_handle=h) for h in handles] if transitions is not None else None   # owners of the handles
        need_e = ctx.needs_input_grad[0]
        need_t = transitions is not None and ctx.needs_input_grad[4]
        with torch.cuda.device(dev):
            sc = rt.host_values_to_device(scales, dev)
            if transitions is not None:
                # alignments := intersect(transitions, alignments) (transducer.py:279-281), with or
                # without epsilon arcs in the transition graph: composition, epsilon fold and the
                # arrays that tie the composed arcs to transition_params for the whole batch in
                # one call of the host library (an epsilon-free graph folds to itself)
                loss, g_e, g_tp = _forward_with_epsilon_transitions(
                    e, handles, transitions, transition_params, sc, need_e, need_t,
                    batch_key=(tokens, lexicon, flat, offs))
                ctx.grads = (g_e if need_e else None, g_tp)
                ctx.devices = (inputs.device, transition_params.device)
                return loss if inputs.is_cuda else loss.cpu()
            packed = _packed_batch(tokens, lexicon, flat, offs, handles, B, dev)
            # loss_b = -(Z_align - Z_norm) * scale_b; mean over b (transducer.py:283-309)
            gs = -sc / B
            z_align, g_e, g_w = lattice_forward_backward(
                e, packed, grad_scale=gs, want_grad_emissions=need_e, want_grad_weights=False)
            score = z_align
            g_tp = None
            loss = (-score * sc).mean()
        ctx.grads = (g_e if need_e else None, g_tp)
        ctx.devices = (inputs.device, transition_params.device if transition_params is not None else None)
        return loss if inputs.is_cuda else loss.cpu()

    @staticmethod
    def backward(ctx, grad_output):
        if ctx.grads is None:
            raise RuntimeError("TransducerLoss: backward called twice on the same forward (the gradient buffers are single "
                               "use; retain_graph is not supported, as in the reference)")
        g_e, g_tp = ctx.grads
        ctx.grads = None
        if g_e is not None:
            g_e = rt.scale_by(g_e, grad_output)
            if g_e.device != ctx.devices[0]:
                g_e = g_e.to(ctx.devices[0])
        if g_tp is not None:
            g_tp = rt.scale_by(g_tp, grad_output)
            if g_tp.device != ctx.devices[1]:
                g_tp = g_tp.to(ctx.devices[1])
        return g_e, None, None, None, g_tp, None, None


TransducerLoss = TransducerLossFunction.apply


class Transducer(torch.nn.Module):
    """A generic transducer loss function (transducer.py:126-197).

    tokens: list of iterables (letters, word pieces, words ...) naming the model outputs;
    graphemes_to_idx: grapheme -> index; ngram: order of a dense token-level transition
    model (0 = none); transitions: a transition Graph (instead of ngram); blank: 'none' |
    'optional' | 'forced'; allow_repeats: if False consecutive equal tokens are not
    allowed in an alignment (needs blank='optional')."""

    def __init__(self, tokens, graphemes_to_idx, ngram=0, transitions=None, blank="none",
                 allow_repeats=True, reduction="none"):
        super().__init__()
        if blank not in ["optional", "forced", "none"]:
            raise ValueError("Invalid value specificed for blank. Must be in ['optional', 'forced', 'none']")
        self.tokens = make_token_graph(tokens, blank=blank, allow_repeats=allow_repeats)
        self.lexicon = make_lexicon_graph(tokens, graphemes_to_idx)
        self.ngram = ngram
        if ngram > 0 and transitions is not None:
            raise ValueError("Only one of ngram and transitions may be specified")
        if ngram > 0:
            transitions = make_transitions_graph(ngram, len(tokens) + int(blank != "none"), True)
        if transitions is not None:
            self.transitions = transitions
            self.transitions.arc_sort()
            self.transition_params = torch.nn.Parameter(torch.zeros(self.transitions.num_arcs()))
        else:
            self.transitions = None
            self.transition_params = None
        self.reduction = reduction

    def forward(self, inputs, targets):
        if self.transitions is None:
            inputs = torch.nn.functional.log_softmax(inputs, dim=2)
        self.tokens.arc_sort(True)
        return TransducerLoss(inputs, targets, self.tokens, self.lexicon, self.transition_params,
                              self.transitions, self.reduction)

    def folded_transitions(self, device):
        """None when the transition graph is epsilon-free, else its folded form (epsilon.py)."""
        if self.transitions is None or not _graph_has_epsilon(self.transitions):
            return None
        return _folded_shared(self.transitions, device)

    def viterbi(self, outputs):
        from ..decode import transducer_viterbi
        return transducer_viterbi(self, outputs)


# --------------------------------------------------------------------------- ConvTransduce1D
def make_kernel_graph(x, blank_idx, blank_optional, spike=False, calc_grad=False):
    """Kernel acceptor of one lexicon entry (criterions/transducer.py:351-367): start in
    blank, each sub-token a label state (with self loop unless `spike`) followed by a blank
    state; with `blank_optional` the blank between different sub-tokens may be skipped."""
    g = G.Graph(calc_grad)
    g.add_node(True, len(x) == 0)
    g.add_arc(0, 0, blank_idx)
    for i, c in enumerate(x):
        g.add_node(False, blank_optional and (i + 1) == len(x))
        g.add_node(False, (i + 1) == len(x))
        g.add_arc(2 * i, 2 * i + 1, c)
        if not spike:
            g.add_arc(2 * i + 1, 2 * i + 1, c)
        g.add_arc(2 * i + 1, 2 * i + 2, blank_idx)
        g.add_arc(2 * i + 2, 2 * i + 2, blank_idx)
        if i > 0 and blank_optional and x[i - 1] != c:
            g.add_arc(2 * i - 1, 2 * i + 1, c)
    g.arc_sort(True)
    g.arc_sort()
    return g


def _packed_kernel(k, dev):
    key = "_packed_%s" % str(dev)
    hit = getattr(k, key, None)
    if hit is None:
        hit = G.pack_graphs([k], dev)
        try:
            setattr(k, key, hit)
        except AttributeError:
            pass
    return hit


def _packed_all_kernels(kernels, dev):
    """all kernel graphs as ONE packed batch (for the single-launch entry), cached on the first graph"""
    if not kernels:
        return None
    key = "_packed_all_%s" % str(dev)
    hit = getattr(kernels[0], key, None)
    ids = [id(k) for k in kernels]
    if hit is None or hit[0] != ids:
        hit = (ids, G.pack_graphs(list(kernels), dev))
        try:
            setattr(kernels[0], key, hit)
        except AttributeError:
            pass
    return hit[1]


def _score_all_kernels(windows, packed, weights, grad_scale=None, grad_emissions=None, grad_weights=None,
                       narcs=None, packed_all=None, flat_weights=None):
    """scores [K, B'] of every window against every kernel graph (wfst_lattice_forward_backward_many);
    with grad_scale [K, B']: window gradients are added to `grad_emissions`, arc-weight gradients
    written to the slices of the flat buffer `grad_weights` (narcs: arcs per kernel graph)."""
    import ctypes
    Bw, ks, C = windows.shape
    K = len(packed)
    dev = windows.device
    L = _lib.lib()
    scores = torch.empty(K, Bw, dtype=torch.float32, device=dev)
    if K == 0:
        return scores
    if packed_all is not None and (flat_weights is not None or all(w is None for w in weights)):
        # one launch: item (k, w) = kernel graph k against window w (wfst_lattice_forward_backward_cross);
        # the library refuses (and launches nothing) when the graphs do not fit the shared-memory kernel
        s = packed_all.struct(flat_weights)
        with torch.cuda.device(dev):
            ws = rt.workspace(dev, L.wfst_lattice_workspace_bytes(K * Bw, ks, C, 0, packed_all.max_nodes))
            rc = L.wfst_lattice_forward_backward_cross(
                windows.data_ptr(), Bw, ks, C, ctypes.byref(s),
                grad_scale.data_ptr() if grad_scale is not None else None, scores.data_ptr(),
                grad_emissions.data_ptr() if grad_emissions is not None else None,
                grad_weights.data_ptr() if grad_weights is not None else None,
                ws.data_ptr(), ws.numel(), rt.stream_ptr(dev))
        if rc == 0:
            return scores
        if rc != -3:
            _lib.check(rc)
    # the K structs of a set of kernel graphs are built once and kept on the first packed graph
    # (the packed graphs themselves are cached on the kernel graphs); only the weight pointers
    # change from call to call
    hit = getattr(packed[0], "_many_structs", None)
    if hit is None or hit[0] != [id(pk) for pk in packed]:
        hit = ([id(pk) for pk in packed], (_lib.AcceptorBatch * K)(*[pk.struct() for pk in packed]),
               [pk.t["weights"].data_ptr() for pk in packed])
        packed[0]._many_structs = hit
    structs = hit[1]
    for i in range(K):
        structs[i].weights = weights[i].data_ptr() if weights[i] is not None else hit[2][i]
    gw_ptrs = None
    if grad_weights is not None:
        gw_ptrs = (ctypes.c_void_p * K)()
        pos = 0
        for i in range(K):
            gw_ptrs[i] = grad_weights.data_ptr() + 4 * pos
            pos += narcs[i]
    with torch.cuda.device(dev):
        ws = rt.workspace(dev, L.wfst_lattice_workspace_bytes(Bw, ks, C, 0, max(pk.max_nodes for pk in packed)))
        _lib.check(L.wfst_lattice_forward_backward_many(
            windows.data_ptr(), Bw, ks, C, structs, K,
            grad_scale.data_ptr() if grad_scale is not None else None, scores.data_ptr(),
            grad_emissions.data_ptr() if grad_emissions is not None else None, gw_ptrs,
            ws.data_ptr(), ws.numel(), rt.stream_ptr(dev)))
    return scores


class ConvTransduce1DFunction(torch.autograd.Function):
    """criterions/transducer.py:461-556.  The reference scores every window of every
    utterance against every kernel graph with one GTN intersect + forward_score (or
    viterbi_score) on CPU threads and keeps all graphs alive in a module-level global for
    the backward pass.  Here the windows become one batch [B*T', kernel_size, C] on the
    device and each kernel graph is one launch of the lattice kernel in shared-graph mode, all
    of them issued by one call of the library (wfst_lattice_forward_backward_many);
    backward re-runs the launches with grad_scale = the incoming output gradient, which
    yields the window gradients (overlap-added into the input gradient) and, summed over
    windows, the kernel weight gradients.  Nothing is kept between calls but the windows."""

    @staticmethod
    def forward(ctx, inputs, kernels, kernel_size, stride, kernel_params=None, viterbi=False):
        from ..decode import lattice_viterbi
        B, T, C = inputs.shape
        if T < kernel_size:
            # Padding should be done outside of this function:
            raise ValueError(f"Input ({T}) too short for kernel ({kernel_size})")
        rt.require_cuda(inputs, "inputs")
        e = rt.to_device(inputs.detach())
        dev = e.device
        with torch.cuda.device(dev):
            win = e.unfold(1, kernel_size, stride)                    # [B, T', C, ks]
            Tp = win.shape[1]
            windows = win.permute(0, 1, 3, 2).reshape(B * Tp, kernel_size, C).contiguous()
            packed = [_packed_kernel(k, dev) for k in kernels]     # packed once per kernel graph and device
            weights = [None] * len(kernels)
            kp = None
            packed_all = _packed_all_kernels(kernels, dev) if not viterbi else None
            if kernel_params is not None:
                kp = kernel_params.detach().to(dev, torch.float32).contiguous()
                pos = 0
                for i, k in enumerate(kernels):
                    weights[i] = kp[pos:pos + k.num_arcs()]
                    pos += k.num_arcs()
            paths = []
            if viterbi:
                out = torch.empty(B * Tp, len(kernels), dtype=torch.float32, device=dev)
                for i, pk in enumerate(packed):
                    sc, labels, arcs = lattice_viterbi(windows, pk, shared=True, weights=weights[i])
                    paths.append((labels, arcs))
                    out[:, i] = sc
            else:
                # every kernel graph against every window in ONE call of the library (the launches
                # are issued back to back from C: no Python work per lexicon entry)
                out = _score_all_kernels(windows, packed, weights, packed_all=packed_all,
                                         flat_weights=kp).t().contiguous()
        ctx.saved = (windows, packed, weights, paths, [k.num_arcs() for k in kernels], packed_all, kp)
        ctx.meta = (B, T, C, Tp, kernel_size, stride, viterbi, inputs.device,
                    kernel_params.device if kernel_params is not None else None)
        out = out.view(B, Tp, len(kernels))
        return out if inputs.is_cuda else out.cpu()

    @staticmethod
    def backward(ctx, grad_output):
        windows, packed, weights, paths, narcs, packed_all, kp = ctx.saved
        B, T, C, Tp, ks, stride, viterbi, in_dev, kp_dev = ctx.meta
        ctx.saved = None
        dev = windows.device
        need_in, need_k = ctx.needs_input_grad[0], ctx.needs_input_grad[4]
        with torch.cuda.device(dev):
            deltas = grad_output.detach().to(dev, torch.float32).reshape(B * Tp, len(packed))
            gwin = torch.zeros_like(windows)
            kgrads = []
            for i, pk in enumerate(packed if viterbi else []):
                gs = deltas[:, i].contiguous()
                # d viterbi_score / d weights = indicator of the best path's arcs
                # (a window with no accepting path has score -inf and labels / arcs of -1:
                # it contributes nothing — the indices are clamped and the values masked)
                labels, arcs = paths[i]
                valid = (labels >= 0).to(torch.float32)                 # [B*T', ks]
                vals = gs.view(-1, 1) * valid
                if need_in:
                    gwin.scatter_add_(2, labels.clamp_min(0).long().unsqueeze(2), vals.unsqueeze(2).contiguous())
                if need_k:
                    kg = torch.zeros(narcs[i], dtype=torch.float32, device=dev)
                    if narcs[i] > 0:
                        kg.index_add_(0, arcs.clamp_min(0).long().reshape(-1), vals.reshape(-1))
                    kgrads.append(kg)
            if not viterbi and len(packed):
                # one call for all kernel graphs: window gradients accumulate in gwin, the weight
                # gradients of kernel graph i land in their slice of one flat buffer
                flat = torch.zeros(sum(narcs), dtype=torch.float32, device=dev) if need_k else None
                _score_all_kernels(windows, packed, weights, grad_scale=deltas.t().contiguous(),
                                   grad_emissions=gwin if need_in else None, grad_weights=flat, narcs=narcs,
                                   packed_all=packed_all, flat_weights=kp)
                if need_k:
                    kgrads = [flat]
            g_in = None
            if need_in:
                g_in = torch.zeros(B, T, C, dtype=torch.float32, device=dev)
                gw = gwin.view(B, Tp, ks, C)
                for j in range(ks):     # window t covers frames t*stride + j
                    g_in[:, j:j + stride * (Tp - 1) + 1:stride] += gw[:, :, j]
                if g_in.device != in_dev:
                    g_in = g_in.to(in_dev)
            g_k = None
            if need_k:
                g_k = torch.cat(kgrads) if kgrads else torch.zeros(0, device=dev)
                if kp_dev is not None and g_k.device != kp_dev:
                    g_k = g_k.to(kp_dev)
        return g_in, None, None, None, g_k, None


class ConvTransduce1D(torch.nn.Module):
    """A 1D convolutional transducer layer (criterions/transducer.py:370-455): same
    constructor, parameters (`kernel_params`) and forward as the reference."""

    def __init__(self, lexicon, kernel_size, stride, blank_idx, blank_optional=True, learn_params=False,
                 scale="none", normalize="none", viterbi=False, spike=False):
        super().__init__()
        import math
        self.normalize = normalize
        self.viterbi = viterbi
        if scale == "none":
            self.scale = 1.0
        elif scale == "sqrt":
            self.scale = math.sqrt(kernel_size)
        elif scale == "linear":
            self.scale = kernel_size
        else:
            raise ValueError(f"Unknown scale {scale}")
        if normalize not in ["none", "pre", "post"]:
            raise ValueError(f"Unknown normalization {normalize}")
        self.kernel_size = kernel_size
        assert self.kernel_size % 2 != 0, "Use an odd kernel size for easy padding."
        self.stride = stride

        def size_with_rep(token):
            reps = sum(t1 == t2 for t1, t2 in zip(token[:-1], token[1:]))
            return len(token) + reps

        min_kernel_size = max(size_with_rep(l) for l in lexicon)
        if kernel_size < min_kernel_size:
            raise ValueError(f"Kernel size needed of at least {min_kernel_size}.")
        self.kernels = [make_kernel_graph(l, blank_idx, blank_optional, spike=spike) for l in lexicon]
        num_arcs = sum(k.num_arcs() for k in self.kernels)
        self.kernel_params = None
        if learn_params:
            self.kernel_params = torch.nn.Parameter(torch.zeros(num_arcs))

    def forward(self, inputs):
        # inputs are of shape [B, T, C]
        pad = self.kernel_size // 2
        inputs = torch.nn.functional.pad(inputs, (0, 0, pad, pad))
        if self.normalize == "pre":
            inputs = torch.nn.functional.log_softmax(inputs, dim=2)
        outputs = ConvTransduce1DFunction.apply(inputs, self.kernels, self.kernel_size, self.stride,
                                                self.kernel_params, self.viterbi)
        outputs = outputs / self.scale
        if self.normalize == "post":
            outputs = torch.nn.functional.softmax(outputs, dim=2)
        if self.normalize == "pre":
            outputs = outputs.exp()
        return outputs
