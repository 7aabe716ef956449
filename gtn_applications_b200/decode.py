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


def asg_viterbi_paths(outputs, transitions):
    """Raw best label path per utterance through emissions o transitions (asg.py:217-226)."""
    from .criterions.asg import ASGLossFunction
    rt.require_cuda(outputs, "outputs")
    e = rt.to_device(outputs.detach())
    g = ASGLossFunction.create_transitions_graph(transitions.detach())
    packed = G.pack_graphs([g], e.device)
    _, labels, _ = lattice_viterbi(e, packed, shared=True)
    return labels.cpu().tolist()


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
        paths = labels.cpu().tolist()
    else:
        paths = torch.argmax(e, dim=2).cpu().tolist()
    crit.tokens.arc_sort()
    preds = []
    for labs in paths:
        chain = G.Graph(False)
        chain.add_node(True, len(labs) == 0)
        for i, lab in enumerate(labs):
            chain.add_node(False, i == len(labs) - 1)
            chain.add_arc(i, i + 1, int(lab))
        best = G.viterbi_path(G.compose(chain, crit.tokens))
        out = G.remove(G.project_output(best))
        preds.append(torch.IntTensor(out.labels_to_list()))
    return preds
