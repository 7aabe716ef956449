"""cfg3 ASG through the Function: time per step, kernels launched (torch profiler), gradient check
of a slice against the float64 oracle (tests/test_gpu_asg.py has the real tests)."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gtn_applications_b200.criterions.asg import ASGLoss
torch.manual_seed(0)
B, T, C, L = 256, 1000, 30, 176
e = torch.randn(B, T, C, device="cuda", requires_grad=True)
tr = torch.randn(C + 1, C, device="cuda", requires_grad=True)
tg = torch.randint(C, (B, L)).tolist()
def asg():
    e.grad = None; tr.grad = None
    ASGLoss(e, tr, tg, "mean").backward()
for _ in range(3): asg()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): asg()
torch.cuda.synchronize()
print("cfg3 ASG: %.3f ms/step" % ((time.perf_counter() - t0) / 10 * 1e3))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    asg(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=8, max_name_column_width=70))
