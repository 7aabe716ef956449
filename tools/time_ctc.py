"""Event-times the CTC C-ABI call at cfg2 (or B T C L from argv) on cuda:0."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gtn_applications_b200 import _lib, _runtime as rt
B, T, C, L = (int(x) for x in sys.argv[1:5]) if len(sys.argv) > 4 else (256, 1000, 30, 176)
iters = 30
torch.manual_seed(0)
dev = torch.device("cuda:0")
lps = [torch.log_softmax(torch.randn(B, T, C, device=dev), 2) for _ in range(4)]
tg = torch.randint(C - 2, (B, L)).tolist()
flat, offsets, _, max_len = rt.pack_targets(tg, C, dev)
gs = torch.full((B,), 1.0 / B, device=dev)
out = torch.empty(B + 1, device=dev); grad = torch.empty_like(lps[0])
Lb = _lib.lib()
Lb.wfst_debug_force_generic_ctc(int(os.environ.get('WFST_CTC_HOOK', '0')))
ws = rt.workspace(dev, Lb.wfst_ctc_workspace_bytes(B, T, C, max_len))
def call(i):
    _lib.check(Lb.wfst_ctc_forward_backward(lps[i % 4].data_ptr(), flat.data_ptr(), offsets.data_ptr(), B, T, C, C - 1,
        max_len, gs.data_ptr(), out.data_ptr(), out[B:].data_ptr(), grad.data_ptr(), ws.data_ptr(), ws.numel(), rt.stream_ptr(dev)))
for i in range(5): call(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(iters): call(i)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print("B=%d T=%d C=%d L=%d: %.4f ms per call (all kernels), %.0f utt/s, %.1f GB/s algorithmic" % (B, T, C, L, ms, B / ms * 1e3, 8.0 * B * T * C / ms / 1e6))
if Lb.wfst_ctc_logits_supported(B, T, C, max_len):
    xs = [torch.randn(B, T, C, device=dev) for _ in range(4)]
    ws2 = rt.workspace(dev, Lb.wfst_ctc_logits_workspace_bytes(B, T, C, max_len))
    def call2(i):
        _lib.check(Lb.wfst_ctc_logits_forward_backward(xs[i % 4].data_ptr(), flat.data_ptr(), offsets.data_ptr(), B, T, C, C - 1,
            max_len, gs.data_ptr(), out.data_ptr(), out[B:].data_ptr(), grad.data_ptr(), ws2.data_ptr(), ws2.numel(), rt.stream_ptr(dev)))
    for i in range(5): call2(i)
    torch.cuda.synchronize()
    e0.record()
    for i in range(iters): call2(i)
    e1.record(); torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / iters
    def two(i):
        lp = torch.log_softmax(xs[i % 4], 2)
        _lib.check(Lb.wfst_ctc_forward_backward(lp.data_ptr(), flat.data_ptr(), offsets.data_ptr(), B, T, C, C - 1,
            max_len, gs.data_ptr(), out.data_ptr(), out[B:].data_ptr(), grad.data_ptr(), ws2.data_ptr(), ws2.numel(), rt.stream_ptr(dev)))
        s = grad.sum(2, keepdim=True)
        g = grad - torch.exp(lp) * s          # what autograd's log_softmax backward does
        return g
    for i in range(5): two(i)
    torch.cuda.synchronize()
    e0.record()
    for i in range(iters): two(i)
    e1.record(); torch.cuda.synchronize()
    ms3 = e0.elapsed_time(e1) / iters
    print("logits path: fused %.4f ms per call; log_softmax + CTC + softmax backward (torch ops) %.4f ms" % (ms2, ms3))
