"""cProfile of the n-gram transducer step's host side (Transducer(ngram=2), B=32, T=250, 81 tokens, L=44)."""
import cProfile, os, pstats, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gtn_applications_b200.criterions.transducer import Transducer
Nn, Tn, Ln, Bn = 81, 250, 44, 32
g = torch.Generator().manual_seed(0)
toks = [(i,) for i in range(Nn)]
gi = {i: i for i in range(Nn)}
xn = torch.randn(Bn, Tn, Nn, generator=g).cuda().requires_grad_(True)
tgn = [t.squeeze() for t in torch.randint(Nn, size=(Bn, Ln), generator=g).split(1)]
cn = Transducer(toks, gi, ngram=2, blank="optional", allow_repeats=False, reduction="mean").cuda()
def step():
    xn.grad = None
    cn(xn, tgn).backward()
for _ in range(3): step()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(10): step()
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
