"""Where the host time of the Function paths goes (GPU box): cProfile of
CTCLoss / ASGLoss with list-of-lists targets and of the e2e-style call with device targets."""
import cProfile, os, pstats, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gtn_applications_b200.criterions.ctc import CTCLoss
from gtn_applications_b200.criterions.asg import ASGLoss
from gtn_applications_b200 import _runtime as rt
torch.manual_seed(0)
B, T, C, L = 256, 1000, 30, 176
dev = torch.device("cuda:0")
lps = [torch.log_softmax(torch.randn(B, T, C, device=dev), 2).requires_grad_(True) for _ in range(5)]
tgt = torch.randint(C - 2, (B, L))
tg = [t.tolist()[0] for t in tgt.split(1)]
tgd = tgt.to(torch.int32).to(dev)
tr = torch.randn(C + 1, C, device=dev).requires_grad_(True)
it = [0]

def ctc_list():
    x = lps[it[0] % 5]; it[0] += 1
    x.grad = None
    CTCLoss(x, tg, C - 1).backward()

def ctc_dev():
    x = lps[it[0] % 5]; it[0] += 1
    x.grad = None
    CTCLoss(x, tgd, C - 1, "none").backward()

def asg_list():
    x = lps[it[0] % 5]; it[0] += 1
    x.grad = None; tr.grad = None
    ASGLoss(x, tr, tg, "mean").backward()

def pack_only():
    rt.pack_targets(tg, C, dev, [1.0 / B] * B)

for name, fn, n in (("pack_targets(list of lists)", pack_only, 200), ("CTCLoss list-of-lists", ctc_list, 200),
                    ("CTCLoss device targets", ctc_dev, 200), ("ASGLoss list-of-lists", asg_list, 50)):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("%-30s host %.3f ms/iter, with sync %.3f ms/iter" % (name, (t1 - t0) / n * 1e3, (t2 - t0) / n * 1e3), flush=True)
    pr = cProfile.Profile(); pr.enable()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(12)
