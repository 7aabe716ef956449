"""Wall-clock (CUDA-synchronised) timing of the other BASELINE configs through the Functions:
cfg3 ASG B=256 T=1000 C=30 L=176; cfg4 transducer B=64 T=1000 1000 word pieces (pieces from a
synthetic list when the reference's benchmarks/word_pieces_tokens_1000.txt is not on the box)."""
import os, random, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gtn_applications_b200.criterions.asg import ASGLoss
from gtn_applications_b200.criterions.transducer import Transducer

def timed(fn, n):
    for _ in range(2): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n

torch.manual_seed(0)
B, T, C, L = 256, 1000, 30, 176
e = torch.randn(B, T, C, device="cuda", requires_grad=True)
tr = torch.randn(C + 1, C, device="cuda", requires_grad=True)
tg = torch.randint(C, (B, L)).tolist()
def asg():
    e.grad = None; tr.grad = None
    ASGLoss(e, tr, tg, "mean").backward()
s = timed(asg, 5)
print("cfg3 ASG B=%d T=%d C=%d L=%d: %.2f ms/step, %.0f utt/s" % (B, T, C, L, s * 1e3, B / s), flush=True)

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
def tdc():
    x.grad = None
    crit(x, targets).backward()
s = timed(tdc, 2)
print("cfg4 transducer B=%d T=%d tokens=%d pieces/utt=%d: %.1f ms/step, %.1f utt/s" % (B, T, Ct, NP, s * 1e3, B / s), flush=True)
