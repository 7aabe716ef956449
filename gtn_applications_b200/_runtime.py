"""Device-side plumbing shared by the criterion Functions: input checks, ragged
target packing, the grow-only workspace and the current-stream handle.  torch is
used for device memory and streams only."""
import torch

from . import _lib

_workspaces = {}


def require_cuda(t, name):
    if t.dtype != torch.float32:
        raise TypeError("%s must be float32 (got %s); the reference reads raw float32 "
                        "through data_ptr() (criterions/ctc.py:43-44)" % (name, t.dtype))
    if not torch.cuda.is_available():
        raise _lib.WfstError(
            "gtn_applications_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    _lib.lib()


def to_device(t):
    """Contiguous float32 CUDA view of `t` (moved if it lives on the host)."""
    if not t.is_cuda:
        t = t.cuda()
    return t.contiguous()


def workspace(device, nbytes):
    buf = _workspaces.get(device)
    if buf is None or buf.numel() < nbytes:
        buf = None
        _workspaces.pop(device, None)
        buf = torch.empty(int(nbytes * 1.25) + 256, dtype=torch.uint8, device=device)
        _workspaces[device] = buf
    return buf


def stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream


def pack_targets(targets, num_classes, device):
    """list[list[int]] -> (flat int32, offsets int32 [B+1], lengths list, max_len) on `device`.
    Labels are validated on the host (the kernels index emissions with them)."""
    lengths = [len(t) for t in targets]
    flat = [int(x) for t in targets for x in t]
    if flat and (min(flat) < 0 or max(flat) >= num_classes):
        raise ValueError("target label outside [0, %d)" % num_classes)
    offsets = [0]
    for n in lengths:
        offsets.append(offsets[-1] + n)
    host = torch.tensor(flat + offsets, dtype=torch.int32)
    dev = host.to(device, non_blocking=False)
    n = len(flat)
    return dev[:n], dev[n:], lengths, (max(lengths) if lengths else 0)


def reduction_scales(reduction, sizes):
    """scale_b = 1/n_b (n_b > 0) for "mean", 1 for "none" (ctc.py:53-58, asg.py:116-121)."""
    if reduction == "mean":
        return [1.0 / n if n > 0 else 1.0 for n in sizes]
    if reduction != "none":
        raise ValueError("invalid value for reduction '" + str(reduction) + "'")
    return [1.0] * len(sizes)


def scale_by(grad, grad_output):
    """grad *= grad_output in place on the device (no pass over grad when grad_output == 1)"""
    go = grad_output.detach().to(device=grad.device, dtype=torch.float32).reshape(1)
    with torch.cuda.device(grad.device):
        _lib.check(_lib.lib().wfst_scale_inplace(grad.data_ptr(), grad.numel(), go.data_ptr(),
                                                 stream_ptr(grad.device)))
    return grad
