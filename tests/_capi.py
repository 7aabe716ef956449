"""Test helper: the CTC hot path called straight through the C ABI
(include/wfst_b200.h: wfst_ctc_forward_backward / wfst_ctc_logits_forward_backward /
wfst_debug_ctc_hazards) on device tensors, returning what the ABI returns
(per-utterance losses, their mean, the [B, T, C] gradient, the fallback flags)."""
import numpy as np
import torch


def ctc_capi(e, targets, blank, scales=None, logits=False):
    """e: [B, T, C] float32 cuda tensor; targets: list of lists.
    Returns (losses [B] numpy, mean float, grad numpy [B, T, C], hazard flags [B])."""
    from gtn_applications_b200 import _lib, _runtime as rt
    L_ = _lib.lib()
    B, T, C = e.shape
    dev = e.device
    lens = [len(t) for t in targets]
    max_len = max(lens) if lens else 0
    flat = torch.tensor([x for t in targets for x in t], dtype=torch.int32, device=dev)
    off = torch.tensor(np.concatenate(([0], np.cumsum(lens))), dtype=torch.int32, device=dev)
    gs = torch.tensor(scales if scales is not None else [1.0 / B] * B, dtype=torch.float32, device=dev)
    out = torch.empty(B + 1, dtype=torch.float32, device=dev)
    grad = torch.empty_like(e)
    if logits:
        assert L_.wfst_ctc_logits_supported(B, T, C, max_len)
        nbytes = L_.wfst_ctc_logits_workspace_bytes(B, T, C, max_len)
        fn = L_.wfst_ctc_logits_forward_backward
    else:
        nbytes = L_.wfst_ctc_workspace_bytes(B, T, C, max_len)
        fn = L_.wfst_ctc_forward_backward
    ws = rt.workspace(dev, nbytes)
    _lib.check(fn(e.data_ptr(), flat.data_ptr(), off.data_ptr(), B, T, C, int(blank), max_len,
                  gs.data_ptr(), out.data_ptr(), out[B:].data_ptr(), grad.data_ptr(), ws.data_ptr(),
                  ws.numel(), torch.cuda.current_stream(dev).cuda_stream))
    torch.cuda.synchronize(dev)
    flags = np.zeros(B, dtype=np.int32)
    _lib.check(L_.wfst_debug_ctc_hazards(ws.data_ptr(), B, T, C, max_len, flags.ctypes.data))
    o = out.cpu().numpy()
    return o[:B].astype(np.float64), float(o[B]), grad.cpu().numpy(), flags
