"""cfg4 lattice kernel alone (B=64, T=1000, the reference's 1000 word pieces, 150 pieces per
utterance): alignment graphs built and packed once, CUDA events around lattice_forward_backward.
Under ncu: `ncu --set full --import-source on -k regex:lattice_lean -c 1 python tools/time_tdc_kernel.py 1`."""
import ctypes, os, random, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import reference_word_pieces
from gtn_applications_b200 import _lib, graph as G, _runtime as rt
from gtn_applications_b200.criterions.transducer import Transducer
from gtn_applications_b200.lattice import lattice_forward_backward

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
tokens, g2i = reference_word_pieces()
rnd = random.Random(0)
Bt, NP, T = 64, 150, 1000
crit = Transducer(tokens, g2i, blank="optional", allow_repeats=False, reduction="mean")
Ct = len(tokens) + 1
torch.manual_seed(0)
x = torch.randn(Bt, T, Ct, device="cuda")
targets = [torch.tensor([g2i[l] for wp in (rnd.choice(tokens) for _ in range(NP)) for l in wp]) for _ in range(Bt)]
flat, offs = rt.flatten_targets_host(targets)
L_ = _lib.lib()
handles = (ctypes.c_int32 * Bt)()
_lib.check(L_.wfst_transducer_alignment_graphs(crit.tokens._h, crit.lexicon._h, flat.ctypes.data, offs.ctypes.data, Bt, handles))
packed = G.pack_handles(handles, Bt, x.device)
print("cfg4 alignment graphs: max nodes %d, max arcs %d" % (packed.max_nodes, packed.max_arcs))
gs = torch.full((Bt,), -1.0 / Bt, device="cuda")
def tdc():
    return lattice_forward_backward(x, packed, grad_scale=gs, want_grad_emissions=True, want_grad_weights=False)
z, ge, _ = tdc()
torch.cuda.synchronize()
print("scores[:4]", z[:4].tolist(), "grad row sums", ge[0, :3].sum(1).tolist(), "checksum %.9e" % ge.double().abs().sum().item())
if n > 1:
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): tdc()
    b.record(); torch.cuda.synchronize()
    print("cfg4 lattice kernel (reference tokens): %.3f ms" % (a.elapsed_time(b) / n))
