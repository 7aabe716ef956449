"""Host side of the reference benchmark's call, CTCLoss(inputs, list_of_lists, blank).backward()
(benchmarks/ctc_benchmark.py:23-29) at cfg2 with device-resident emissions: cProfile + wall clock."""
import cProfile, os, pstats, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gtn_applications_b200.criterions.ctc import CTCLoss
torch.manual_seed(0)
B, T, C, L = 256, 1000, 30, 176
lps = [torch.log_softmax(torch.randn(B, T, C, device="cuda"), 2).requires_grad_(True) for _ in range(5)]
tg = torch.randint(C - 1, (B, L)).tolist()
it = [0]
def step():
    x = lps[it[0] % 5]; it[0] += 1
    x.grad = None
    CTCLoss(x, tg, C - 1).backward()
for _ in range(5): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(100): step()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("host loop %.3f ms per call, with final synchronise %.3f ms per call" % ((t1 - t0) * 10, (t2 - t0) * 10))
pr = cProfile.Profile(); pr.enable()
for _ in range(100): step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(16)
