"""CTC on B200 — same interface as the reference's criterions/ctc.py.

``CTCLossFunction.forward(ctx, log_probs, targets, blank_idx=0, reduction="none")``
and ``.backward(ctx, grad_output) -> (grad, None, None, None)`` mirror
criterions/ctc.py:32,72,89-94; ``CTCLoss = CTCLossFunction.apply`` (ctc.py:97);
``CTC(blank, use_pt)`` mirrors ctc.py:100-135.  Where the reference builds one
GTN graph pair per utterance on CPU threads and copies every utterance to the
host (ctc.py:40-51), this launches one fused forward+backward kernel for the
whole batch on the current CUDA stream (wfst_ctc_forward_backward) and keeps
only the [B,T,C] gradient for ``backward``.
"""
import torch

from .. import _lib, _runtime as rt


class CTCLossFunction(torch.autograd.Function):
    @staticmethod
    def create_ctc_graph(target, blank_idx):
        """The CTC acceptor of criterions/ctc.py:15-29 as a host Graph (the CUDA
        kernel derives the same chain in closed form; this exists for API parity
        and for tests that inspect the graph)."""
        from ..graph import Graph
        g = Graph(False)
        n_states = 2 * len(target) + 1
        for s in range(n_states):
            k = (s - 1) // 2
            g.add_node(s == 0, s >= n_states - 2)
            lab = target[k] if s % 2 else blank_idx
            g.add_arc(s, s, lab)
            if s > 0:
                g.add_arc(s - 1, s, lab)
            if s % 2 and s > 1 and lab != target[k - 1]:
                g.add_arc(s - 2, s, lab)
        g.arc_sort(False)
        return g

    @staticmethod
    def forward(ctx, log_probs, targets, blank_idx=0, reduction="none"):
        if log_probs.dim() != 3:
            raise ValueError("log_probs must be [B, T, C]")
        B, T, C = log_probs.shape
        rt.require_cuda(log_probs, "log_probs")
        if reduction not in ("none", "mean"):
            raise ValueError("invalid value for reduction '" + str(reduction) + "'")
        if len(targets) != B:
            raise ValueError("need one target sequence per batch entry")
        if not 0 <= blank_idx < C:
            raise ValueError("blank_idx outside [0, C)")
        e = rt.to_device(log_probs.detach())
        dev = e.device
        with torch.cuda.device(dev):
            flat, offsets, max_len, gscale = rt.pack_targets_reduction(targets, C, dev, reduction, B)
            out = torch.empty(B + 1, dtype=torch.float32, device=dev)
            need_grad = log_probs.requires_grad
            grad = torch.empty_like(e) if need_grad else None
            L = _lib.lib()
            nbytes = L.wfst_ctc_workspace_bytes(B, T, C, max_len)
            ws = rt.workspace(dev, nbytes)
            _lib.check(L.wfst_ctc_forward_backward(
                e.data_ptr(), flat.data_ptr(), offsets.data_ptr(), B, T, C, int(blank_idx),
                max_len, gscale.data_ptr(), out.data_ptr(), out[B:].data_ptr(),
                grad.data_ptr() if need_grad else None, ws.data_ptr(), ws.numel(),
                rt.stream_ptr(dev)))
        ctx.grad = grad
        ctx.input_device = log_probs.device
        ctx.per_utterance_loss = out[:B]
        loss = out[B]
        return loss if log_probs.is_cuda else loss.cpu()

    @staticmethod
    def backward(ctx, grad_output):
        if getattr(ctx, "consumed", False):
            # gtn.backward(..., retain_graph=False) frees the tape (ctc.py:78): a second backward
            # through the same forward is an error there too (skipped test gtn_ctc_test.py:82)
            raise RuntimeError("CTCLoss: backward called twice on the same forward (the gradient buffer is "
                               "single use; retain_graph is not supported, as in the reference)")
        grad = ctx.grad
        ctx.grad = None
        ctx.consumed = True
        if grad is None:
            return None, None, None, None
        go = grad_output.detach().to(device=grad.device, dtype=torch.float32).reshape(1)
        with torch.cuda.device(grad.device):
            _lib.check(_lib.lib().wfst_scale_inplace(
                grad.data_ptr(), grad.numel(), go.data_ptr(), rt.stream_ptr(grad.device)))
        if grad.device != ctx.input_device:
            grad = grad.to(ctx.input_device)
        return grad, None, None, None


CTCLoss = CTCLossFunction.apply


class CTCLogitsLossFunction(torch.autograd.Function):
    """``CTCLoss(log_softmax(inputs, 2), targets, blank_idx, reduction)`` — what
    ``CTC.forward`` computes (criterions/ctc.py:107 followed by ctc.py:31-94) — as ONE
    kernel on raw logits (wfst_ctc_logits_forward_backward): the log-softmax is applied
    to the tiles as they are staged, and the gradient that comes back is already
    d loss / d inputs = scale * (softmax - posterior), the form the reference's golden
    tests check (tests/gtn_ctc_test.py:64-80).  Shapes the fused kernel does not
    handle (``supported`` is False) must use log_softmax + CTCLoss."""

    @staticmethod
    def supported(inputs, targets):
        if not (inputs.is_cuda and inputs.dtype == torch.float32 and inputs.dim() == 3):
            return False
        B, T, C = inputs.shape
        if inputs.is_contiguous() and inputs.data_ptr() % 16 != 0:
            return False        # the fused kernel moves tiles with 16-byte bulk copies
        max_len = max(rt.target_lengths(targets), default=0)
        return bool(_lib.lib().wfst_ctc_logits_supported(B, T, C, max_len))

    @staticmethod
    def forward(ctx, inputs, targets, blank_idx=0, reduction="none"):
        if inputs.dim() != 3:
            raise ValueError("inputs must be [B, T, C]")
        B, T, C = inputs.shape
        rt.require_cuda(inputs, "inputs")
        if reduction not in ("none", "mean"):
            raise ValueError("invalid value for reduction '" + str(reduction) + "'")
        if len(targets) != B:
            raise ValueError("need one target sequence per batch entry")
        if not 0 <= blank_idx < C:
            raise ValueError("blank_idx outside [0, C)")
        e = rt.to_device(inputs.detach())
        dev = e.device
        with torch.cuda.device(dev):
            flat, offsets, max_len, gscale = rt.pack_targets_reduction(targets, C, dev, reduction, B)
            L = _lib.lib()
            if not L.wfst_ctc_logits_supported(B, T, C, max_len):
                raise NotImplementedError("fused logits CTC does not handle this shape; use log_softmax + CTCLoss")
            out = torch.empty(B + 1, dtype=torch.float32, device=dev)
            need_grad = inputs.requires_grad
            grad = torch.empty_like(e) if need_grad else None
            ws = rt.workspace(dev, L.wfst_ctc_logits_workspace_bytes(B, T, C, max_len))
            _lib.check(L.wfst_ctc_logits_forward_backward(
                e.data_ptr(), flat.data_ptr(), offsets.data_ptr(), B, T, C, int(blank_idx),
                max_len, gscale.data_ptr(), out.data_ptr(), out[B:].data_ptr(),
                grad.data_ptr() if need_grad else None, ws.data_ptr(), ws.numel(),
                rt.stream_ptr(dev)))
        ctx.grad = grad
        ctx.input_device = inputs.device
        ctx.per_utterance_loss = out[:B]
        loss = out[B]
        return loss if inputs.is_cuda else loss.cpu()

    backward = staticmethod(CTCLossFunction.backward)


CTCLogitsLoss = CTCLogitsLossFunction.apply


class CTC(torch.nn.Module):
    """criterions/ctc.py:100-135.  `use_pt` selects torch's ctc_loss exactly as
    the reference does; otherwise the B200 kernel is used."""

    def __init__(self, blank, use_pt):
        super().__init__()
        self.blank = blank
        self.use_pt = use_pt

    def forward(self, inputs, targets):
        if not self.use_pt and CTCLogitsLossFunction.supported(inputs, targets):
            # log_softmax fused into the kernel (same value and gradient as the two steps below);
            # the label tensors are packed without a detour through Python lists
            return CTCLogitsLoss(inputs, list(targets), self.blank, "mean")
        log_probs = torch.nn.functional.log_softmax(inputs, dim=2)
        if self.use_pt:
            lengths = [t.numel() for t in targets]
            return torch.nn.functional.ctc_loss(
                log_probs.permute(1, 0, 2), torch.cat(targets), [inputs.shape[1]] * inputs.shape[0],
                lengths, blank=self.blank, zero_infinity=True)
        return CTCLoss(log_probs, [t.tolist() for t in targets], self.blank, "mean")

    def viterbi(self, outputs):
        """Greedy decode: argmax per frame, collapse repeats, drop blanks (ctc.py:126-135)."""
        best = torch.argmax(outputs, dim=2)
        keep = torch.ones_like(best, dtype=torch.bool)
        keep[:, 1:] = best[:, 1:] != best[:, :-1]
        keep &= best != self.blank
        best, keep = best.cpu(), keep.cpu()
        return [best[b][keep[b]] for b in range(best.shape[0])]
