"""ASG on B200 — same interface as the reference's criterions/asg.py.

``ASGLossFunction.forward(ctx, inputs, transitions, targets, reduction="none")`` /
``.backward -> (input_grad, transitions_grad, None, None)`` mirror asg.py:84,142,
180-185; ``ASG(num_classes, num_replabels=1, use_garbage=True)`` keeps the
``transitions`` parameter name and [(N+1), N] shape (asg.py:192-199), which is
part of the checkpoint format (train.py:117)."""
import itertools

import torch

from .. import _lib, _runtime as rt


def pack_replabels(tokens, num_replabels):
    """Replace runs of a repeated token by the token followed by a repetition
    label (asg.py:13-32): labels 0..num_replabels-1 mean "repeat 1..n more
    times", real tokens are shifted up by num_replabels."""
    if all(isinstance(t, list) for t in tokens):
        return [pack_replabels(t, num_replabels) for t in tokens]
    assert isinstance(tokens, list)
    out = []
    pending = 0
    last = -1
    for tok in tokens:
        if tok == last and pending < num_replabels:
            pending += 1
            continue
        if pending > 0:
            out.append(pending - 1)
            pending = 0
        out.append(tok + num_replabels)
        last = tok
    if pending > 0:
        out.append(pending - 1)
    return out


def unpack_replabels(tokens, num_replabels):
    """Inverse of pack_replabels (asg.py:35-49)."""
    if all(isinstance(t, list) for t in tokens):
        return [unpack_replabels(t, num_replabels) for t in tokens]
    assert isinstance(tokens, list)
    out = []
    last = -1
    for tok in tokens:
        if tok >= num_replabels:
            out.append(tok - num_replabels)
            last = tok
        elif last != -1:
            out.extend([last - num_replabels] * (tok + 1))
            last = -1
    return out


def unpack_replabels_batch(flat, counts, num_replabels):
    """unpack_replabels for a whole batch at once: `flat` holds the token sequences of B
    utterances back to back (`counts[b]` tokens each).  A token >= num_replabels is a label; a
    replabel token r directly after a label repeats that label r + 1 times; any other replabel
    is dropped (asg.py:35-49).  Returns a list of B int32 tensors."""
    import numpy as np
    flat = np.asarray(flat, dtype=np.int64)
    counts = np.asarray(counts, dtype=np.int64)
    B, R = len(counts), num_replabels
    uid = np.repeat(np.arange(B), counts)
    prev_same = np.zeros(len(flat), dtype=bool)
    prev_same[1:] = uid[1:] == uid[:-1]
    prev_tok = np.zeros(len(flat), dtype=np.int64)
    prev_tok[1:] = flat[:-1]
    is_label = flat >= R
    eff_rep = (~is_label) & prev_same & (prev_tok >= R)
    reps = np.where(is_label, 1, np.where(eff_rep, flat + 1, 0))
    vals = np.where(is_label, flat - R, prev_tok - R)
    out = torch.from_numpy(np.repeat(vals, reps).astype(np.int32))
    sizes = np.bincount(np.repeat(uid, reps), minlength=B)
    return list(torch.split(out, sizes.tolist()))


class ASGLossFunction(torch.autograd.Function):
    @staticmethod
    def create_transitions_graph(transitions, calc_grad=False):
        """Bigram acceptor of asg.py:53-69 as a host Graph (API parity; the ASG
        kernels read the [(C+1), C] matrix directly)."""
        from ..graph import Graph
        C = transitions.shape[1]
        assert transitions.shape == (C + 1, C)
        g = Graph(calc_grad)
        g.add_node(True)
        for i in range(1, C + 1):
            g.add_node(False, True)
            g.add_arc(0, i, i - 1)
        for i in range(C):
            for j in range(C):
                g.add_arc(j + 1, i + 1, i)
        g.set_weights(transitions.detach().cpu().contiguous().reshape(-1).tolist())
        g.mark_arc_sorted(False)
        g.mark_arc_sorted(True)
        return g

    @staticmethod
    def create_force_align_graph(target):
        """asg.py:71-81."""
        from ..graph import Graph
        g = Graph(False)
        g.add_node(True)
        for k in range(1, len(target) + 1):
            g.add_node(False, k == len(target))
            g.add_arc(k - 1, k, target[k - 1])
            g.add_arc(k, k, target[k - 1])
        g.arc_sort(True)
        return g

    @staticmethod
    def forward(ctx, inputs, transitions, targets, reduction="none"):
        B, T, C = inputs.shape
        rt.require_cuda(inputs, "inputs")
        rt.require_cuda(transitions, "transitions")
        assert transitions.shape == (C + 1, C)
        if reduction not in ("none", "mean"):
            raise ValueError("invalid value for reduction '" + str(reduction) + "'")
        e = rt.to_device(inputs.detach())
        dev = e.device
        with torch.cuda.device(dev):
            tr = transitions.detach().to(dev).contiguous()
            flat, offsets, max_len, gscale = rt.pack_targets_reduction(targets, C, dev, reduction, B)
            out = torch.empty(B + 1, dtype=torch.float32, device=dev)
            need_e, need_t = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
            g_e = torch.empty_like(e) if need_e else None
            g_t = torch.empty_like(tr) if need_t else None
            L = _lib.lib()
            ws = rt.workspace(dev, L.wfst_asg_workspace_bytes(B, T, C, max_len))
            _lib.check(L.wfst_asg_forward_backward(
                e.data_ptr(), tr.data_ptr(), flat.data_ptr(), offsets.data_ptr(), B, T, C, max_len,
                gscale.data_ptr(), out.data_ptr(), out[B:].data_ptr(),
                g_e.data_ptr() if g_e is not None else None,
                g_t.data_ptr() if g_t is not None else None,
                ws.data_ptr(), ws.numel(), rt.stream_ptr(dev)))
        ctx.grads = (g_e, g_t)
        ctx.devices = (inputs.device, transitions.device)
        loss = out[B]
        return loss if inputs.is_cuda else loss.cpu()

    @staticmethod
    def backward(ctx, grad_output):
        if ctx.grads is None:
            raise RuntimeError("ASGLoss: backward called twice on the same forward (the gradient buffers are single "
                               "use; retain_graph is not supported, as in the reference)")
        g_e, g_t = ctx.grads
        ctx.grads = None
        L = _lib.lib()
        for g in (g_e, g_t):
            if g is not None:
                go = grad_output.detach().to(device=g.device, dtype=torch.float32).reshape(1)
                with torch.cuda.device(g.device):
                    _lib.check(L.wfst_scale_inplace(g.data_ptr(), g.numel(), go.data_ptr(),
                                                    rt.stream_ptr(g.device)))
        if g_e is not None and g_e.device != ctx.devices[0]:
            g_e = g_e.to(ctx.devices[0])
        if g_t is not None and g_t.device != ctx.devices[1]:
            g_t = g_t.to(ctx.devices[1])
        return g_e, g_t, None, None


ASGLoss = ASGLossFunction.apply


class ASG(torch.nn.Module):
    """asg.py:191-237."""

    def __init__(self, num_classes, num_replabels=1, use_garbage=True):
        super().__init__()
        self.num_classes = num_classes
        self.num_replabels = num_replabels
        assert self.num_replabels > 0
        self.garbage_idx = (num_classes + num_replabels) if use_garbage else None
        self.N = num_classes + num_replabels + int(use_garbage)
        self.transitions = torch.nn.Parameter(torch.zeros(self.N + 1, self.N))

    def forward(self, inputs, targets):
        packed = [pack_replabels(t.tolist(), self.num_replabels) for t in targets]
        if self.garbage_idx is not None:
            # a garbage token before, between and after the labels (asg.py:203-208)
            with_garbage = []
            for p in packed:
                seq = [self.garbage_idx] * (2 * len(p) + 1)
                seq[1::2] = p
                with_garbage.append(seq)
            packed = with_garbage
        return ASGLoss(inputs, self.transitions, packed, "mean")

    def viterbi(self, outputs):
        from ..decode import asg_viterbi_collapsed
        B, T, C = outputs.shape
        assert C == self.N, "Wrong number of classes in output."
        # repeats merged, then garbage dropped (asg.py:228-233), on the device for the whole batch
        flat, counts = asg_viterbi_collapsed(outputs, self.transitions, self.garbage_idx)
        return unpack_replabels_batch(flat, counts, self.num_replabels)
