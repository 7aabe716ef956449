"""ORACLE — TEST INFRASTRUCTURE ONLY.

Second, independent oracle: the closed-form time-synchronous recursions of
SURVEY.md Appendix B in float64 numpy — no graph composition at all.  For an
epsilon-free acceptor A with arcs (src, dst, label, w):

    alpha_0[q]     = 0 if q is a start node else -inf
    alpha_{t+1}[v] = LSE_{(u->v, c, w)} alpha_t[u] + E[t, c] + w
    Z              = LSE_{q accept} alpha_T[q]
    dZ/dE[t, c]    = sum_{arcs with label c} exp(alpha_t[u] + E[t,c] + w + beta_{t+1}[v] - Z)
    dZ/dw_a        = sum_t  of the same term

which is what forward_score(intersect(linear_graph(T, C), A)) followed by
gtn.backward computes (ctc.py:49-51,78; asg.py:111-115,158; stc.py:85-86,113;
transducer.py:283-290,321).  Used to cross-check the GTN restatement
(tests/test_oracle_dp.py) and as float64 truth at sizes the materialising
oracle is too slow for.
"""
import numpy as np

NEG = -np.inf


def _lse_scatter(values, index, size):
    """out[i] = logsumexp(values[index == i]); -inf where empty."""
    m = np.full(size, NEG)
    np.maximum.at(m, index, values)
    safe = np.where(np.isfinite(m), m, 0.0)
    s = np.zeros(size)
    with np.errstate(invalid="ignore"):
        np.add.at(s, index, np.exp(values - safe[index]))
    with np.errstate(divide="ignore"):
        out = safe + np.log(s)
    return np.where(np.isfinite(m), out, NEG)


def acceptor_forward_backward(E, start, accept, src, dst, label, weight, want_grad=True):
    """E [T, C] float; start/accept boolean [N]; arcs as int / float arrays.
    Returns Z and (if want_grad) dZ/dE [T, C], dZ/dweight [A]."""
    E = np.asarray(E, dtype=np.float64)
    T, C = E.shape
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    label = np.asarray(label, dtype=np.int64)
    weight = np.asarray(weight, dtype=np.float64)
    N = len(start)
    alpha = np.full((T + 1, N), NEG)
    alpha[0, np.asarray(start, dtype=bool)] = 0.0
    for t in range(T):
        x = alpha[t, src] + E[t, label] + weight
        alpha[t + 1] = _lse_scatter(x, dst, N)
    acc = np.asarray(accept, dtype=bool)
    fin = alpha[T, acc]
    if fin.size == 0 or not np.isfinite(fin.max()):
        Z = NEG
    else:
        Z = fin.max() + np.log(np.exp(fin - fin.max()).sum())
    if not want_grad:
        return Z
    gE = np.zeros((T, C))
    gW = np.zeros(len(src))
    if not np.isfinite(Z):
        return Z, gE, gW
    beta = np.full(N, NEG)
    beta[acc] = 0.0
    for t in range(T - 1, -1, -1):
        x = E[t, label] + weight + beta[dst]
        with np.errstate(invalid="ignore"):
            post = np.exp(alpha[t, src] + x - Z)
        post = np.where(np.isfinite(alpha[t, src]) & np.isfinite(x), post, 0.0)
        np.add.at(gE[t], label, post)
        gW += post
        beta = _lse_scatter(x, src, N)
    return Z, gE, gW


# ---- closed-form criterion acceptors (no gtn involved) --------------------
def ctc_acceptor(target, blank):
    """CTC states s in [0, 2L]; in-arcs of s: s, s-1, and s-2 when s is odd
    and the label differs from the previous one (Appendix B)."""
    L = len(target)
    S = 2 * L + 1
    src, dst, lab = [], [], []
    for s in range(S):
        k = (s - 1) // 2
        l = target[k] if s % 2 else blank
        src.append(s); dst.append(s); lab.append(l)
        if s > 0:
            src.append(s - 1); dst.append(s); lab.append(l)
        if s % 2 and s > 1 and l != target[k - 1]:
            src.append(s - 2); dst.append(s); lab.append(l)
    start = np.zeros(S, dtype=bool); start[0] = True
    accept = np.zeros(S, dtype=bool); accept[S - 1] = True
    if S >= 2:
        accept[S - 2] = True
    return start, accept, np.array(src), np.array(dst), np.array(lab), np.zeros(len(src))


def ctc(E, targets, blank, reduction="none"):
    """Batch-mean CTC loss and its gradient w.r.t. E (ctc.py:31-94)."""
    E = np.asarray(E, dtype=np.float64)
    B = E.shape[0]
    losses = np.zeros(B)
    grad = np.zeros_like(E)
    for b in range(B):
        Z, gE, _ = acceptor_forward_backward(E[b], *ctc_acceptor(list(targets[b]), blank))
        L = len(targets[b])
        scale = (1.0 / L if L > 0 else 1.0) if reduction == "mean" else 1.0
        losses[b] = -Z * scale
        grad[b] = -gE * scale / B
    return {"loss": losses.mean(), "losses": losses, "grad": grad}


def asg(E, transitions, targets, reduction="none"):
    """ASG closed form (Appendix B): FAL over target positions, FCC over the
    C classes; transitions[0, i] = start score, transitions[1+i, j] = i | j."""
    E = np.asarray(E, dtype=np.float64)
    tr = np.asarray(transitions, dtype=np.float64)
    B, T, C = E.shape
    losses = np.zeros(B)
    gE_all = np.zeros_like(E)
    gT_all = np.zeros((B,) + tr.shape)
    # full-connect acceptor: node 0 start, nodes 1..C accept
    f_src = [0] * C + [j + 1 for i in range(C) for j in range(C)]
    f_dst = list(range(1, C + 1)) + [i + 1 for i in range(C) for j in range(C)]
    f_lab = list(range(C)) + [i for i in range(C) for j in range(C)]
    f_start = np.zeros(C + 1, dtype=bool); f_start[0] = True
    f_acc = ~f_start
    for b in range(B):
        y = list(targets[b])
        L = len(y)
        scale = (1.0 / L if L > 0 else 1.0) if reduction == "mean" else 1.0
        Zc, gEc, gWc = acceptor_forward_backward(E[b], f_start, f_acc, f_src, f_dst, f_lab, tr.reshape(-1))
        # force-align acceptor: nodes 0..L; (l-1 -> l) and (l -> l), label y_l,
        # weight = transition into y_l from the previous label (or <s>)
        a_src, a_dst, a_lab, a_tix = [], [], [], []
        for l in range(1, L + 1):
            cur = y[l - 1]
            a_src.append(l - 1); a_dst.append(l); a_lab.append(cur)
            a_tix.append(cur if l == 1 else C + cur * C + y[l - 2])
            a_src.append(l); a_dst.append(l); a_lab.append(cur)
            a_tix.append(C + cur * C + cur)
        a_start = np.zeros(L + 1, dtype=bool); a_start[0] = True
        a_acc = np.zeros(L + 1, dtype=bool); a_acc[L] = True
        a_tix = np.array(a_tix, dtype=np.int64)
        Za, gEa, gWa = acceptor_forward_backward(
            E[b], a_start, a_acc, a_src, a_dst, a_lab, tr.reshape(-1)[a_tix] if L else np.zeros(0))
        losses[b] = (Zc - Za) * scale
        gE_all[b] = (gEc - gEa) * scale / B
        gt = gWc.copy()
        np.subtract.at(gt, a_tix, gWa)
        gT_all[b] = gt.reshape(tr.shape) * scale
    return {"loss": losses.mean(), "losses": losses, "grad": gE_all,
            "grad_transitions": gT_all.mean(0)}


# ---- CTC without an arc list: the three-term recursion over states (Appendix B) ----
def ctc_dense_one(E, target, blank):
    """One utterance: E [T, C] float64 -> (log Z, dZ/dE [T, C]).  Same recursion as
    acceptor_forward_backward(E, *ctc_acceptor(target, blank)), vectorised over the 2L+1
    states (shifted copies instead of an arc list), so that full BASELINE-size batches
    (256 x T=1000 x S=353) finish in seconds."""
    E = np.asarray(E, dtype=np.float64)
    T, C = E.shape
    y = np.asarray(list(target), dtype=np.int64)
    L = len(y)
    S = 2 * L + 1
    lab = np.full(S, blank, dtype=np.int64)
    lab[1::2] = y
    skip = np.zeros(S, dtype=bool)
    if L > 1:
        skip[3::2] = y[1:] != y[:-1]
    if T == 0:
        return (0.0 if L == 0 else NEG), np.zeros((T, C))
    Es = E[:, lab]                                            # [T, S]
    def shift_right(v, k):      # out[s] = v[s - k]
        out = np.full(S, NEG)
        if k < S:
            out[k:] = v[:S - k]
        return out

    def shift_left(v, k):       # out[s] = v[s + k]
        out = np.full(S, NEG)
        if k < S:
            out[:S - k] = v[k:]
        return out

    def lse3(a, b, c):
        m = np.maximum(np.maximum(a, b), c)
        safe = np.where(np.isfinite(m), m, 0.0)
        with np.errstate(divide="ignore", invalid="ignore"):
            out = safe + np.log(np.exp(a - safe) + np.exp(b - safe) + np.exp(c - safe))
        return np.where(np.isfinite(m), out, NEG)

    alpha = np.full((T, S), NEG)
    alpha[0, 0] = Es[0, 0]
    if S > 1:
        alpha[0, 1] = Es[0, 1]
    for t in range(1, T):
        a = alpha[t - 1]
        a1 = shift_right(a, 1)
        a2 = np.where(skip, shift_right(a, 2), NEG)
        alpha[t] = lse3(a, a1, a2) + Es[t]
    fin = alpha[T - 1, max(S - 2, 0):]
    if not np.isfinite(fin.max()):
        return NEG, np.zeros((T, C))
    Z = fin.max() + np.log(np.exp(fin - fin.max()).sum())
    # beta[t, s]: log mass of completing from state s after frame t (emission of t excluded)
    beta = np.full((T, S), NEG)
    beta[T - 1, max(S - 2, 0):] = 0.0
    skip_out = np.zeros(S, dtype=bool)                                # arc s -> s+2 exists
    if S > 2:
        skip_out[:S - 2] = skip[2:]
    for t in range(T - 2, -1, -1):
        x = beta[t + 1] + Es[t + 1]
        x1 = shift_left(x, 1)
        x2 = np.where(skip_out, shift_left(x, 2), NEG)
        beta[t] = lse3(x, x1, x2)
    with np.errstate(invalid="ignore"):
        post = np.exp(alpha + beta - Z)
    post = np.where(np.isfinite(alpha) & np.isfinite(beta), post, 0.0)
    onehot = np.zeros((S, C))
    onehot[np.arange(S), lab] = 1.0
    return Z, post @ onehot


def ctc_dense(E, targets, blank, reduction="none"):
    """ctc() on the vectorised recursion (same return value)."""
    E = np.asarray(E, dtype=np.float64)
    B = E.shape[0]
    losses = np.zeros(B)
    grad = np.zeros_like(E)
    for b in range(B):
        Z, gE = ctc_dense_one(E[b], targets[b], blank)
        L = len(targets[b])
        scale = (1.0 / L if L > 0 else 1.0) if reduction == "mean" else 1.0
        losses[b] = -Z * scale
        grad[b] = -gE * scale / B
    return {"loss": losses.mean(), "losses": losses, "grad": grad}
