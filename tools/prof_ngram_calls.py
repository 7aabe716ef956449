"""Wall time of every call into libwfst_b200.so and the GPU kernels during n-gram transducer steps
(Transducer(ngram=2), B=32, T=250, 81 tokens, L=44: the shapes of transducer_benchmark.py:56-119)."""
import collections, os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
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
_lib.lib = lambda p=Proxy(): p
Nn, Tn, Ln, Bn = 81, 250, 44, 32
g = torch.Generator().manual_seed(0)
toks = [(i,) for i in range(Nn)]
gi = {i: i for i in range(Nn)}
xn = torch.randn(Bn, Tn, Nn, generator=g).cuda().requires_grad_(True)
tgn = [t.squeeze() for t in torch.randint(Nn, size=(Bn, Ln), generator=g).split(1)]
for name, kw in (("ngram_ctc", dict(ngram=2, blank="optional", allow_repeats=False, reduction="mean")),
                 ("ngram_asg", dict(ngram=2, reduction="mean"))):
    cn = Transducer(toks, gi, **kw).cuda()
    def step():
        xn.grad = None
        cn(xn, tgn).backward()
    for _ in range(3): step()
    torch.cuda.synchronize(); acc.clear()
    n = 10
    t0 = time.perf_counter()
    for _ in range(n): step()
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("%s: host loop %.2f ms per step, with final synchronise %.2f ms per step" % (name, (t1 - t0) / n * 1e3, (t2 - t0) / n * 1e3))
    for k, (c, t) in sorted(acc.items(), key=lambda kv: -kv[1][1])[:8]:
        print("   %-40s %4d calls  %.3f ms per step" % (k, c, t / n * 1e3))
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step(); torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=6, max_name_column_width=60))
