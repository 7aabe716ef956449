import cProfile, os, pstats, random, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gtn_applications_b200.criterions.transducer import Transducer
random.seed(0)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pieces = sorted([l.strip() for l in open(os.path.join(ROOT, "tests", "golden", "word_pieces_tokens_1000.txt"))])
letters = sorted(set(c for t in pieces for c in t))
g2i = {ch: i for i, ch in enumerate(letters)}
B, T, NP = 64, 1000, 150
crit = Transducer(pieces, g2i, blank="optional", allow_repeats=False, reduction="mean")
Ct = len(pieces) + 1
x = torch.randn(B, T, Ct, device="cuda", requires_grad=True)
targets = [torch.tensor([g2i[c] for c in "".join(random.choice(pieces) for _ in range(NP))]) for _ in range(B)]
def step():
    x.grad = None
    crit(x, targets).backward()
for _ in range(3): step()
torch.cuda.synchronize()
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
for _ in range(6): step()
torch.cuda.synchronize()
pr.disable()
print("per step %.2f ms" % ((time.perf_counter() - t0) / 6 * 1e3))
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
# GPU time of one step
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record(); step(); e1.record(); torch.cuda.synchronize()
print("one step, GPU events: %.2f ms" % e0.elapsed_time(e1))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=8))
