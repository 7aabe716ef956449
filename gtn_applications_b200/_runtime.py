"""Device-side plumbing shared by the criterion Functions: input checks, ragged
target packing, the grow-only workspace and the current-stream handle.  torch is
used for device memory and streams only."""
import torch

from . import _lib

_workspaces = {}
_rect_offsets = {}   # (device, B, L) -> int32 offsets of a rectangular [B, L] target tensor
_rect_scales = {}    # (device, B, L, scale) -> float32 [B] vector of equal scales


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
    """Grow-only scratch buffer of the CURRENT STREAM of `device`.  The kernels of one criterion
    call use it from launch to completion (alpha history, checkpoints, fallback flags), and calls
    on one stream are ordered, so one buffer per (device, stream) is what makes criteria that run
    concurrently on different streams (a side-stream eval, DataParallel threads, the ASG side
    stream) not alias each other's scratch."""
    key = (str(device), int(torch.cuda.current_stream(device).cuda_stream))
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = None
        _workspaces.pop(key, None)
        if len(_workspaces) > 16:       # streams come and go: do not keep scratch of dead ones forever
            _workspaces.clear()
        buf = torch.empty(int(nbytes * 1.25) + 256, dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


def host_values_to_device(values, device, dtype=torch.float32):
    """small host list -> device tensor through pinned staging and an asynchronous copy.
    torch.tensor(values, device=...) copies from pageable memory and then synchronises the
    stream: the host would wait for the previous step's kernels there (measured on the cfg4
    transducer step: 1.6 ms of GPU idle time per step)."""
    host = torch.empty(len(values), dtype=dtype, pin_memory=torch.cuda.is_available())
    if len(values):
        host.numpy()[:] = values
    return host.to(device, non_blocking=True)


def stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream


def pack_targets(targets, num_classes, device, scales=None):
    """list[list[int]] -> (flat int32, offsets int32 [B+1], lengths list, max_len) on `device`.
    Labels are validated on the host (the kernels index emissions with them).  With `scales`
    (a list of B floats) a fifth value is returned: the float32 device vector of those scales,
    shipped in the same pinned staging buffer — one asynchronous H2D copy per call, so that a
    criterion call never waits for the stream."""
    import itertools
    import numpy as np
    if torch.is_tensor(targets) and targets.dim() == 2 and targets.is_cuda:
        # extension over the reference: a rectangular label tensor that already lives on the
        # device is used in place — no host round trip and no H2D copy on the compute stream
        # (a small copy there queues behind a loader's bulk H2D copies on the copy engine and
        # delays the kernel).  Labels are not range-checked on the host; the kernels clamp them.
        B, L = targets.shape
        flat = targets.detach().to(device=device, dtype=torch.int32).contiguous().reshape(-1)
        key = (str(device), int(B), int(L))
        offs = _rect_offsets.get(key)
        if offs is None:
            offs = (torch.arange(B + 1, dtype=torch.int32, device=device) * L).contiguous()
            _rect_offsets[key] = offs
        out = (flat, offs, [L] * B, L)
        if scales is not None:
            skey = key + (float(scales[0]) if len(scales) else 0.0,)
            gs = _rect_scales.get(skey)
            if gs is None or any(s != scales[0] for s in scales):
                gs = host_values_to_device(scales, device)
                if all(s == scales[0] for s in scales):
                    if len(_rect_scales) > 64:
                        _rect_scales.clear()
                    _rect_scales[skey] = gs
            out = out + (gs,)
        return out
    if torch.is_tensor(targets) and targets.dim() == 2:
        # extension over the reference (which takes Python lists): a rectangular integer
        # tensor [B, L] skips the per-label Python work
        B, L = targets.shape
        t = targets.detach().to("cpu", torch.int32).contiguous()
        if t.numel() and (int(t.min()) < 0 or int(t.max()) >= num_classes):
            raise ValueError("target label outside [0, %d)" % num_classes)
        lengths, total = [L] * B, B * L
        flat = t.reshape(-1).numpy()
        offs = np.arange(B + 1, dtype=np.int32) * L
    else:
        fast = _pack_list_of_lists(targets, num_classes, device, scales)
        if fast is not None:
            return fast
        lengths = [len(t) for t in targets]
        total = sum(lengths)
        if lengths and all(torch.is_tensor(t) for t in targets):
            # a list of 1-D label tensors (what train.py hands to CTC.forward): one concatenation
            # instead of per-label Python work
            flat = torch.cat([t.detach().reshape(-1) for t in targets]).to("cpu", torch.int64).numpy()
        else:
            flat = np.fromiter(itertools.chain.from_iterable(targets), dtype=np.int64, count=total)
        if total and (flat.min() < 0 or flat.max() >= num_classes):
            raise ValueError("target label outside [0, %d)" % num_classes)
        offs = np.zeros(len(lengths) + 1, dtype=np.int32)
        np.cumsum(lengths, out=offs[1:])
    nb = len(lengths)
    ns = nb if scales is not None else 0
    host = torch.empty(total + nb + 1 + ns, dtype=torch.int32, pin_memory=torch.cuda.is_available())
    buf = host.numpy()
    buf[:total] = flat
    buf[total:total + nb + 1] = offs
    if ns:
        buf[total + nb + 1:].view(np.float32)[:] = np.asarray(scales, dtype=np.float32)
    dev = host.to(device, non_blocking=True)
    out = (dev[:total], dev[total:total + nb + 1], lengths, (max(lengths) if lengths else 0))
    if scales is not None:
        out = out + (dev[total + nb + 1:].view(torch.float32),)
    return out


def pack_targets_reduction(targets, num_classes, device, reduction, batch):
    """pack_targets + reduction_scales in one step: returns (flat, offsets, max_len, gscale) where
    gscale[b] = scale_b / batch (ctc.py:53-56,81,87).  For the reference's list-of-lists argument
    the lengths come out of the C walker and the scales are formed vectorised — no per-utterance
    Python work on the call path (the Function is host-bound at BASELINE sizes otherwise)."""
    import numpy as np
    if reduction not in ("none", "mean"):
        raise ValueError("invalid value for reduction '" + str(reduction) + "'")
    if isinstance(targets, (list, tuple)) and targets and isinstance(targets[0], (list, tuple)):
        fast = _pack_list_of_lists(targets, num_classes, device, None, (reduction, batch))
        if fast is not None:
            return fast
    scales = reduction_scales(reduction, target_lengths(targets))
    flat, offsets, _, max_len, gscale = pack_targets(targets, num_classes, device, [x / batch for x in scales])
    return flat, offsets, max_len, gscale


def _pack_list_of_lists(targets, num_classes, device, scales, reduction=None):
    """The reference's own argument type — a Python list of lists of ints (ctc.py:32,
    benchmarks/ctc_benchmark.py:23-24) — walked in C (csrc/pytargets.c) straight into the pinned
    staging buffer: 0.1 ms instead of 1.6 ms of per-label interpreter work at B=256, L=176.
    Returns None when `targets` is anything else (tensors, numpy scalars, generators)."""
    import numpy as np
    if not isinstance(targets, (list, tuple)) or not targets or not isinstance(targets[0], (list, tuple)):
        return None
    pt = _lib.pytargets()
    nb = len(targets)
    lens = np.empty(nb, dtype=np.int32)
    total = pt.wfst_pytargets_lengths(targets, lens.ctypes.data, nb)
    if total < 0:
        return None
    ns = nb if (scales is not None or reduction is not None) else 0
    host = torch.empty(total + nb + 1 + ns, dtype=torch.int32, pin_memory=torch.cuda.is_available())
    mm = np.zeros(2, dtype=np.int32)
    if pt.wfst_pytargets_fill(targets, host.data_ptr(), total, mm.ctypes.data) != total:
        return None
    if total and (mm[0] < 0 or mm[1] >= num_classes):
        raise ValueError("target label outside [0, %d)" % num_classes)
    buf = host.numpy()
    buf[total] = 0
    np.cumsum(lens, out=buf[total + 1:total + nb + 1])
    if reduction is not None:
        kind, batch = reduction
        sc = buf[total + nb + 1:].view(np.float32)
        if kind == "mean":
            np.divide(1.0 / batch, np.maximum(lens, 1), out=sc)
        else:
            sc[:] = 1.0 / batch
    elif ns:
        buf[total + nb + 1:].view(np.float32)[:] = scales
    dev = host.to(device, non_blocking=True)
    max_len = int(lens.max()) if nb else 0
    if reduction is not None:
        return dev[:total], dev[total:total + nb + 1], max_len, dev[total + nb + 1:].view(torch.float32)
    out = (dev[:total], dev[total:total + nb + 1], lens.tolist(), max_len)
    if scales is not None:
        out = out + (dev[total + nb + 1:].view(torch.float32),)
    return out


def flatten_targets_host(targets):
    """ragged targets (lists, 1-D tensors or a [B, L] tensor) -> (flat int32, offsets int32 [B+1])
    numpy arrays on the host, without per-label Python work for tensors (iterating a tensor
    element by element costs ~1 us per label: 30 ms for a 64 x 400 grapheme batch)"""
    import itertools
    import numpy as np
    if torch.is_tensor(targets) and targets.dim() == 2:
        B, L = targets.shape
        flat = targets.detach().to("cpu", torch.int32).contiguous().reshape(-1).numpy()
        return flat, (np.arange(B + 1, dtype=np.int32) * L)
    lengths = [len(t) for t in targets]
    total = sum(lengths)
    if lengths and all(torch.is_tensor(t) for t in targets):
        flat = torch.cat([t.detach().reshape(-1) for t in targets]).to("cpu", torch.int32).contiguous().numpy() \
            if total else np.zeros(0, dtype=np.int32)
    else:
        flat = np.fromiter(itertools.chain.from_iterable(targets), dtype=np.int32, count=total)
    offs = np.zeros(len(lengths) + 1, dtype=np.int32)
    np.cumsum(lengths, out=offs[1:])
    return np.ascontiguousarray(flat, dtype=np.int32), offs


def target_lengths(targets):
    """lengths of the target sequences; a [B, L] tensor is not iterated row by row"""
    if torch.is_tensor(targets) and targets.dim() == 2:
        return [int(targets.shape[1])] * int(targets.shape[0])
    return [len(t) for t in targets]


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
