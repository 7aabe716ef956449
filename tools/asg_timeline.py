import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gtn_applications_b200 import _lib, _runtime as rt
from torch.profiler import profile, ProfilerActivity
L_ = _lib.lib()
torch.manual_seed(0)
B, T, C, L = 256, 1000, 30, 176
e = torch.randn(B, T, C, device="cuda"); tr = torch.randn(C + 1, C, device="cuda")
tgt = [torch.randint(C, (L,)) for _ in range(B)]
flat_a, offs_a, _, maxlen_a, gsc_a = rt.pack_targets(tgt, C, e.device, [1.0 / (B * L)] * B)
out_a = torch.empty(B + 1, device="cuda"); ge_a = torch.empty_like(e); gt_a = torch.empty_like(tr)
ws_a = rt.workspace(e.device, L_.wfst_asg_workspace_bytes(B, T, C, maxlen_a))
def asg_abi():
    _lib.check(L_.wfst_asg_forward_backward(
        e.data_ptr(), tr.data_ptr(), flat_a.data_ptr(), offs_a.data_ptr(), B, T, C, maxlen_a, gsc_a.data_ptr(),
        out_a.data_ptr(), out_a[B:].data_ptr(), ge_a.data_ptr(), gt_a.data_ptr(), ws_a.data_ptr(), ws_a.numel(),
        torch.cuda.current_stream().cuda_stream))
for _ in range(3): asg_abi()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): asg_abi()
    torch.cuda.synchronize()
evs = sorted([ev for ev in prof.events() if ev.device_type.name == "CUDA"], key=lambda ev: ev.time_range.start)
t0 = evs[0].time_range.start
for ev in evs:
    print("%9.1f us  +%8.1f us  %s" % (ev.time_range.start - t0, ev.time_range.end - ev.time_range.start, ev.name[:70]))
