"""Wall time of every call into libwfst_b200.so during cfg4 transducer steps (B=64, T=1000, the
reference's word pieces; alignment graphs and the packed batch come from the LRU caches)."""
import collections, os, random, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import reference_word_pieces
from gtn_applications_b200 import _lib
from gtn_applications_b200.criterions.transducer import Transducer

real = _lib.lib()
acc = collections.defaultdict(lambda: [0, 0.0])
class Proxy:
    def __getattr__(self, name):
        f = getattr(real, name)
        def g(*a):
            t0 = time.perf_counter()
            r = f(*a)
            e = acc[name]; e[0] += 1; e[1] += time.perf_counter() - t0
            return r
        return g
proxy = Proxy()
_lib.lib = lambda: proxy

tokens, g2i = reference_word_pieces()
rnd = random.Random(0)
B, T, NP = 64, 1000, 150
crit = Transducer(tokens, g2i, blank="optional", allow_repeats=False, reduction="mean")
x = torch.randn(B, T, len(tokens) + 1, device="cuda", requires_grad=True)
targets = [torch.tensor([g2i[l] for wp in (rnd.choice(tokens) for _ in range(NP)) for l in wp]) for _ in range(B)]
def step():
    x.grad = None
    crit(x, targets).backward()
for _ in range(3): step()
torch.cuda.synchronize()
acc.clear()
n = 10
t0 = time.perf_counter()
for _ in range(n): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host loop %.2f ms per step, with final synchronise %.2f ms per step" % ((t1 - t0) / n * 1e3, (t2 - t0) / n * 1e3))
for k, (c, t) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print("%-40s %4d calls  %.3f ms per step" % (k, c, t / n * 1e3))
