"""Star Temporal Classification on B200 — same interface as the reference's
criterions/stc.py.  ``STCLossFunction.forward(ctx, inputs, targets, prob,
reduction="none")`` / ``.backward -> (grad, None, None, None)`` mirror stc.py:67,107,
124-129; ``STC(blank_idx, p0, plast, thalf, reduction)`` mirrors stc.py:135-221.  The
per-utterance STC acceptors are built on the host exactly as create_stc_graph does
(stc.py:22-64) and scored by the generic lattice kernel
(wfst_lattice_forward_backward) instead of gtn.compose + forward_score + backward."""
import math

import torch

from .. import _runtime as rt
from .. import _lib
from ..graph import Graph, pack_graphs, pack_handles, destroy_handles
from ..lattice import lattice_forward_backward

# blank idx is REQUIRED to be zero, as in the reference (stc.py:12)
STC_BLANK_IDX = 0


class STCLossFunction(torch.autograd.Function):
    """Assumes <star>, <star>\\token columns are appended to the input (stc.py:17-19)."""

    @staticmethod
    def create_stc_graph(target, star_idx, prob):
        """stc.py:22-64: CTC-like chain with self loops only on blank states and
        unconditional skip arcs, plus one <star> node per gap whose entering / looping
        arcs carry log(prob)."""
        g = Graph(False)
        L = len(target)
        n_states = 2 * L + 1
        for s in range(n_states):
            k = (s - 1) // 2
            g.add_node(s == 0, s >= n_states - 2)
            lab = target[k] if s % 2 else STC_BLANK_IDX
            if lab == STC_BLANK_IDX:
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
            g.add_arc(c, prev_blank, STC_BLANK_IDX)
        return g

    @staticmethod
    def forward(ctx, inputs, targets, prob, reduction="none"):
        B, T, Cstar = inputs.shape
        rt.require_cuda(inputs, "inputs")
        if reduction == "mean":
            scale = 1.0 / T if T > 0 else 1.0       # stc.py:90-91: "mean" divides by T
        elif reduction == "none":
            scale = 1.0
        else:
            raise ValueError("invalid value for reduction '" + str(reduction) + "'")
        star = Cstar // 2
        e = rt.to_device(inputs.detach())
        dev = e.device
        # the acceptors of the whole batch from the host library (wfst_stc_graphs: create_stc_graph
        # below, same node and arc order, on host threads — no Python work per arc), packed once
        import ctypes
        if len(targets) != B:
            raise ValueError("%d targets for a batch of %d" % (len(targets), B))
        flat, offs = rt.flatten_targets_host([t if torch.is_tensor(t) else list(t) for t in targets])
        handles = (ctypes.c_int32 * B)()
        _lib.check(_lib.lib().wfst_stc_graphs(flat.ctypes.data, offs.ctypes.data, B, star, math.log(prob),
                                              STC_BLANK_IDX, handles))
        with torch.cuda.device(dev):
            try:
                packed = pack_handles(handles, B, dev)
            finally:
                destroy_handles(handles, B)
            gscale = torch.full((B,), -scale / B, dtype=torch.float32, device=dev)
            scores, grad, _ = lattice_forward_backward(
                e, packed, grad_scale=gscale, want_grad_emissions=inputs.requires_grad)
            loss = (-scale * scores).mean()
        ctx.grad = grad
        ctx.input_device = inputs.device
        return loss if inputs.is_cuda else loss.cpu()

    @staticmethod
    def backward(ctx, grad_output):
        if getattr(ctx, "consumed", False):
            raise RuntimeError("STCLoss: backward called twice on the same forward (the gradient buffers are single "
                               "use; retain_graph is not supported, as in the reference)")
        grad = ctx.grad
        ctx.grad = None
        ctx.consumed = True
        if grad is None:
            return None, None, None, None
        grad = rt.scale_by(grad, grad_output)
        if grad.device != ctx.input_device:
            grad = grad.to(ctx.input_device)
        return grad, None, None, None


STCLoss = STCLossFunction.apply


class STC(torch.nn.Module):
    """The Star Temporal Classification loss (stc.py:135-221): loss between an
    unsegmented time series and a partially labelled target.

    p0 / plast: initial / final token insertion penalty (before the log); thalf: number
    of steps after which the penalty is halfway between them."""

    def __init__(self, blank_idx, p0=1, plast=1, thalf=1, reduction="none"):
        super().__init__()
        assert blank_idx == STC_BLANK_IDX
        self.p0 = p0
        self.plast = plast
        self.thalf = thalf
        self.nstep = 0
        self.reduction = reduction

    @staticmethod
    def logsubexp(a, b):
        """log(exp(a) - exp(b)) with a [M,N,1] broadcast against b [M,N,O] (stc.py:157-172)"""
        with torch.set_grad_enabled(a.requires_grad):
            a = a.expand(-1, -1, b.shape[2])
            return a + torch.log1p(1e-7 - torch.exp(b - a))

    def forward(self, inputs, targets):
        """inputs: [T, B, C] log-probabilities; targets: list of label lists.  Returns the
        batch-mean STC loss."""
        if self.training:
            self.nstep += 1
        prob = self.plast + (self.p0 - self.plast) * math.exp(-self.nstep * math.log(2) / self.thalf)
        log_probs = inputs.permute(1, 0, 2)
        with torch.set_grad_enabled(log_probs.requires_grad):
            # <star>: everything but blank
            lse = torch.logsumexp(log_probs[:, :, 1:], 2, keepdim=True)
            # keep only the tokens that occur in this batch (stc.py:204-214)
            present = [STC_BLANK_IDX] + list(set(t for tgt in targets for t in tgt))
            remap = {t: i for i, t in enumerate(present)}
            idx = rt.host_values_to_device(present, log_probs.device, torch.long)
            sel = log_probs.index_select(2, idx)
            targets = [[remap[t] for t in tgt] for tgt in targets]
            # <star>\token for every present token, then [tokens, <star>, <star>\tokens]
            neg = STC.logsubexp(lse, sel[:, :, 1:])
            feats = torch.cat([sel, lse, neg], dim=2)
        return STCLoss(feats, targets, prob, self.reduction)
