"""forward_score(intersect(emissions, A_b)) + gtn.backward for packed epsilon-free
acceptors on the GPU (wfst_lattice_forward_backward)."""
import ctypes

import torch

from . import _lib, _runtime as rt


def lattice_forward_backward(emissions, packed, grad_scale=None, want_grad_emissions=True,
                             want_grad_weights=False, weights=None, shared=False,
                             accumulate_into=None, final_weights=None):
    """emissions [B,T,C] float32 CUDA; packed: PackedAcceptors.  Returns
    (scores [B], grad_emissions or None, grad_weights or None); gradients are
    grad_scale[b] * dZ_b/d(.).  With `final_weights` ([nodes], epsilon.py) a fourth value is
    returned: the gradient w.r.t. the final weights (when want_grad_weights)."""
    B, T, C = emissions.shape
    dev = emissions.device
    L = _lib.lib()
    scores = torch.empty(B, dtype=torch.float32, device=dev)
    if accumulate_into is not None:
        g_e, acc = accumulate_into, 1
    else:
        g_e, acc = (torch.empty_like(emissions) if want_grad_emissions else None), 0
    # an output of the call (the kernels / the C entry point clear what they accumulate into)
    g_w = torch.empty(packed.num_arcs, dtype=torch.float32, device=dev) if want_grad_weights else None
    g_f = None
    if final_weights is not None and want_grad_weights:
        g_f = torch.zeros(final_weights.numel(), dtype=torch.float32, device=dev)
    s = packed.struct(weights, final_weights, g_f)
    with torch.cuda.device(dev):
        ws = rt.workspace(dev, L.wfst_lattice_workspace_bytes(B, T, C, 0, packed.max_nodes))
        _lib.check(L.wfst_lattice_forward_backward(
            emissions.data_ptr(), B, T, C, ctypes.byref(s), 1 if shared else 0,
            grad_scale.data_ptr() if grad_scale is not None else None, scores.data_ptr(),
            g_e.data_ptr() if g_e is not None else None, acc,
            g_w.data_ptr() if g_w is not None else None, ws.data_ptr(), ws.numel(),
            rt.stream_ptr(dev)))
    if final_weights is not None:
        return scores, g_e, g_w, g_f
    return scores, g_e, g_w
