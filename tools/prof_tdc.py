import os, random, sys, time, torch
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
targets = []
for _ in range(B):
    word = "".join(random.choice(pieces) for _ in range(NP))
    targets.append(torch.tensor([g2i[c] for c in word]))
for i in range(2):
    x.grad = None
    torch.cuda.synchronize(); t0 = time.perf_counter()
    loss = crit(x, targets)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    loss.backward(); torch.cuda.synchronize(); t3 = time.perf_counter()
    print("forward host %.1f ms, +gpu wait %.1f ms, backward %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
