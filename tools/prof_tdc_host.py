"""Where the host time of a cfg4 transducer forward goes (B=64, T=1000, 1000 word pieces):
cProfile of Transducer.forward, top entries by cumulative time."""
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
for _ in range(2):
    crit(x, targets).backward(); torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    loss = crit(x, targets)
    torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
print("os.cpu_count()", os.cpu_count())
