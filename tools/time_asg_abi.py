"""cfg3 ASG at the C ABI (CUDA events), as bench.py's asg_abi."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gtn_applications_b200 import _lib, _runtime as rt
L_ = _lib.lib()
B, T, C, L = 256, 1000, 30, 176
dev = torch.device("cuda:0"); torch.manual_seed(0)
es = [torch.randn(B, T, C, device=dev) for _ in range(4)]
tr = torch.randn(C + 1, C, device=dev)
tg = torch.randint(C, (B, L)).tolist()
flat, offsets, _, max_len = rt.pack_targets(tg, C, dev)
gs = torch.full((B,), 1.0 / B, device=dev)
out = torch.empty(B + 1, device=dev); ge = torch.empty_like(es[0]); gt = torch.empty_like(tr)
ws = rt.workspace(dev, L_.wfst_asg_workspace_bytes(B, T, C, max_len))
def call(i):
    _lib.check(L_.wfst_asg_forward_backward(es[i % 4].data_ptr(), tr.data_ptr(), flat.data_ptr(), offsets.data_ptr(), B, T, C, max_len,
        gs.data_ptr(), out.data_ptr(), out[B:].data_ptr(), ge.data_ptr(), gt.data_ptr(), ws.data_ptr(), ws.numel(), rt.stream_ptr(dev)))
for i in range(5): call(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(30): call(i)
e1.record(); torch.cuda.synchronize()
print("ASG cfg3 at the ABI: %.4f ms per call" % (e0.elapsed_time(e1) / 30))
