"""Host-side overhead of the Function paths at cfg2 (wall clock per call incl. sync)."""
import os, sys, time, cProfile, pstats, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gtn_applications_b200.criterions.ctc import CTCLoss, CTCLogitsLoss
B, T, C, L = 256, 1000, 30, 176
x = torch.randn(B, T, C, device="cuda")
tgt = torch.randint(C - 2, (B, L))
for name, tg in (("list-of-lists targets", tgt.tolist()), ("[B,L] tensor targets", tgt)):
    def fused():
        a = x.requires_grad_(True); a.grad = None
        CTCLogitsLoss(a, tg, C - 1, "mean").backward()
    def two():
        a = x.requires_grad_(True); a.grad = None
        CTCLoss(torch.log_softmax(a, 2), tg, C - 1, "mean").backward()
    for fn in (fused, two):
        for _ in range(3): fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(30): fn()
        torch.cuda.synchronize(); print(name, fn.__name__, "%.3f ms/call" % ((time.perf_counter() - t0) / 30 * 1e3))
pr = cProfile.Profile(); pr.enable()
for _ in range(20): fused()
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(8)
