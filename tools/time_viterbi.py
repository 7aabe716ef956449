"""Decode timing at cfg3 (ASG B=256, T=1000, C=30): GPU best path through emissions o
transitions (event-timed kernel path) and the whole ASG.viterbi call (wall clock)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gtn_applications_b200 import graph as G
from gtn_applications_b200.criterions.asg import ASG, ASGLossFunction
from gtn_applications_b200.decode import lattice_viterbi
torch.manual_seed(0)
B, T, C = 256, 1000, 30
e = torch.randn(B, T, C, device="cuda")
tr = torch.randn(C + 1, C, device="cuda")
g = ASGLossFunction.create_transitions_graph(tr)
packed = G.pack_graphs([g], e.device)
from gtn_applications_b200.decode import asg_viterbi_labels
def ev(fn):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / 5
print("ASG best path (B=%d,T=%d,C=%d): generic kernel on the packed transition graph %.3f ms, dense kernel %.3f ms"
      % (B, T, C, ev(lambda: lattice_viterbi(e, packed, shared=True)), ev(lambda: asg_viterbi_labels(e, tr))))
crit = ASG(C - 1, num_replabels=1, use_garbage=False).cuda()
with torch.no_grad(): crit.transitions.copy_(tr[:crit.N + 1, :crit.N] if crit.N != C else tr)
x = e[:, :, :crit.N].contiguous()
for _ in range(2): crit.viterbi(x)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(3): crit.viterbi(x)
torch.cuda.synchronize()
print("ASG.viterbi whole call: %.2f ms" % ((time.perf_counter() - t0) / 3 * 1e3))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(3): crit.viterbi(x)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(8)
