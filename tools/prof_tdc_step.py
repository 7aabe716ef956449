import cProfile, os, pstats, random, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gtn_applications_b200.criterions.transducer import Transducer
random.seed(0)
letters = "abcdefghijklmnopqrstuvwxyz"
pieces = sorted({"".join(random.choice(letters) for _ in range(random.randint(1, 4))) for _ in range(1400)})[:1000]
for ch in letters:
    if ch not in pieces: pieces[random.randrange(len(pieces))] = ch
pieces = sorted(set(pieces))
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
