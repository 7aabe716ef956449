"""Decoding (tropical semiring) on the GPU — the Viterbi paths behind ASG.viterbi
(criterions/asg.py:211-237) and Transducer.viterbi (criterions/transducer.py:199-234)."""
import ctypes

import torch

from . import _lib, _runtime as rt
from . import graph as G


def lattice_viterbi(emissions, packed, shared=False, weights=None, final_weights=None):
    """Best path through emissions o acceptor for every utterance.  Returns (scores [B],
    labels [B,T] int32, arcs [B,T] int32) on the device."""
    B, T, C = emissions.shape
    dev = emissions.device
    L = _lib.lib()
    scores = torch.empty(B, dtype=torch.float32, device=dev)
    labels = torch.empty(B, T, dtype=torch.int32, device=dev)
    arcs = torch.empty(B, T, dtype=torch.int32, device=dev)
    s = packed.struct(weights, final_weights)
    with torch.cuda.device(dev):
        ws = rt.workspace(dev, L.wfst_lattice_viterbi_workspace_bytes(B, T, packed.max_nodes))
        _lib.check(L.wfst_lattice_viterbi(
            emissions.data_ptr(), B, T, C, ctypes.byref(s), 1 if shared else 0, scores.data_ptr(),
            labels.data_ptr(), arcs.data_ptr(), ws.data_ptr(), ws.numel(), rt.stream_ptr(dev)))
    return scores, labels, arcs


def asg_viterbi_labels(outputs, transitions):
    """Best label path [B, T] (int32, on the device) through emissions o transitions
    (asg.py:217-226): the dense warp-per-utterance kernel when the shape fits it, else the
    generic best-path kernel on the packed transition graph (same result, ties included)."""
    from .criterions.asg import ASGLossFunction
    rt.require_cuda(outputs, "outputs")
    e = rt.to_device(outputs.detach())
    B, T, C = e.shape
    L = _lib.lib()
    if L.wfst_asg_viterbi_supported(T, C):
        tr = transitions.detach().to(e.device, torch.float32).contiguous()
        scores = torch.empty(B, dtype=torch.float32, device=e.device)
        labels = torch.empty(B, T, dtype=torch.int32, device=e.device)
        with torch.cuda.device(e.device):
            _lib.check(L.wfst_asg_viterbi(e.data_ptr(), tr.data_ptr(), B, T, C, scores.data_ptr(),
                                          labels.data_ptr(), rt.stream_ptr(e.device)))
        return labels
    g = ASGLossFunction.create_transitions_graph(transitions.detach())
    packed = G.pack_graphs([g], e.device)
    _, labels, _ = lattice_viterbi(e, packed, shared=True)
    return labels


def asg_viterbi_paths(outputs, transitions):
    """Raw best label path per utterance (list of lists)."""
    return asg_viterbi_labels(outputs, transitions).cpu().tolist()


def asg_viterbi_collapsed(outputs, transitions, drop_label=None):
    """Best paths with consecutive repeats merged (and `drop_label` removed afterwards) — the
    itertools.groupby / garbage filter of asg.py:228-233 — done on the device for the whole
    batch; one device -> host copy of the surviving labels.  Returns (flat int32 numpy array of
    all surviving labels, numpy array of the B per-utterance counts)."""
    import numpy as np
    labels = asg_viterbi_labels(outputs, transitions)
    B, T = labels.shape
    if T == 0:
        return np.zeros(0, dtype=np.int32), np.zeros(B, dtype=np.int64)
    keep = torch.ones(B, T, dtype=torch.bool, device=labels.device)
    keep[:, 1:] = labels[:, 1:] != labels[:, :-1]
    if drop_label is not None:
        keep &= labels != drop_label
    counts = keep.sum(1).cpu().numpy()
    flat = labels[keep].cpu().numpy()
    return flat, counts


def transducer_viterbi(crit, outputs):
    """Transducer.viterbi (transducer.py:199-234): best alignment through the emissions
    (composed with the transition graph when there is one) on the GPU, then the
    alignment -> token mapping (compose with the token graph, best path, project, remove
    epsilons) on the host graphs."""
    rt.require_cuda(outputs, "outputs")
    e = rt.to_device(outputs.detach())
    if crit.transitions is not None:
        tp = crit.transition_params.detach().to(e.device, torch.float32).contiguous()
        crit.transitions.calc_grad = False
        folded = crit.folded_transitions(e.device)
        if folded is None:
            packed = G.pack_graphs([crit.transitions], e.device)
            _, labels, _ = lattice_viterbi(e, packed, shared=True, weights=tp)
        else:
            # epsilon arcs (</s>, back-off) folded into arcs / final weights (epsilon.py)
            packed, ties = folded
            w, fw, _ = ties.weights(tp, tropical=True)
            _, labels, _ = lattice_viterbi(e, packed, shared=True, weights=w, final_weights=fw)
        paths = labels
    else:
        paths = torch.argmax(e, dim=2)
    crit.tokens.arc_sort()
    # alignment -> tokens (transducer.py:223-233) for the whole batch on host threads
    lab = paths.to("cpu", torch.int32).contiguous()
    B, T = lab.shape
    out = torch.empty(B, max(T, 1), dtype=torch.int32)
    counts = torch.empty(B, dtype=torch.int32)
    _lib.check(_lib.lib().wfst_transducer_decode_paths(
        crit.tokens._h, lab.data_ptr(), B, T, out.data_ptr(), counts.data_ptr()))
    return [out[b, :n].clone() for b, n in enumerate(counts.tolist())]
