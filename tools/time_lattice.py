"""CUDA-event timing of the acceptor lattice kernels, shared-memory ("lean") kernel against the
generic global-memory kernel (forced through wfst_debug_force_generic_lattice):
  cfg3 ASG step (dense full-connect + force-align), cfg4 transducer alignment lattices
  (graphs packed once, kernel only), CTC log-semiring fallback at cfg2."""
import os, random, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gtn_applications_b200 import _lib, graph as G
from gtn_applications_b200.criterions.asg import ASGLoss
from gtn_applications_b200.criterions.ctc import CTCLoss
from gtn_applications_b200.criterions.transducer import Transducer
from gtn_applications_b200.lattice import lattice_forward_backward
import ctypes, numpy as np

def ev_time(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

L_ = _lib.lib()
torch.manual_seed(0)
B, T, C, L = 256, 1000, 30, 176
e = torch.randn(B, T, C, device="cuda", requires_grad=True)
tr = torch.randn(C + 1, C, device="cuda", requires_grad=True)
tg = torch.randint(C, (B, L)).tolist()
tgt = [torch.tensor(t) for t in tg]
def asg():
    e.grad = None; tr.grad = None
    ASGLoss(e, tr, tgt, "mean").backward()
# the same step at the C ABI (no per-call Python work): targets packed once
from gtn_applications_b200 import _runtime as rt
flat_a, offs_a, _, maxlen_a, gsc_a = rt.pack_targets(tgt, C, e.device, [1.0 / (B * L)] * B)
out_a = torch.empty(B + 1, device="cuda"); ge_a = torch.empty_like(e); gt_a = torch.empty_like(tr)
ws_a = rt.workspace(e.device, L_.wfst_asg_workspace_bytes(B, T, C, maxlen_a))
def asg_abi():
    _lib.check(L_.wfst_asg_forward_backward(
        e.data_ptr(), tr.data_ptr(), flat_a.data_ptr(), offs_a.data_ptr(), B, T, C, maxlen_a, gsc_a.data_ptr(),
        out_a.data_ptr(), out_a[B:].data_ptr(), ge_a.data_ptr(), gt_a.data_ptr(), ws_a.data_ptr(), ws_a.numel(),
        torch.cuda.current_stream().cuda_stream))
lp = torch.log_softmax(torch.randn(B, T, C, device="cuda"), 2).requires_grad_(True)
tgc = [torch.randint(C - 1, (L,)) for _ in range(B)]
def ctc():
    lp.grad = None
    CTCLoss(lp, tgc, C - 1, "mean").backward()

random.seed(0)
letters = "abcdefghijklmnopqrstuvwxyz"
pieces = sorted({"".join(random.choice(letters) for _ in range(random.randint(1, 4))) for _ in range(1400)})[:1000]
for ch in letters:
    if ch not in pieces: pieces[random.randrange(len(pieces))] = ch
pieces = sorted(set(pieces))
g2i = {ch: i for i, ch in enumerate(letters)}
Bt, NP = 64, 150
crit = Transducer(pieces, g2i, blank="optional", allow_repeats=False, reduction="mean")
Ct = len(pieces) + 1
x = torch.randn(Bt, T, Ct, device="cuda")
targets = [[g2i[c] for c in "".join(random.choice(pieces) for _ in range(NP))] for _ in range(Bt)]
flat = np.ascontiguousarray([v for t in targets for v in t], dtype=np.int32)
offs = np.zeros(Bt + 1, dtype=np.int32); offs[1:] = np.cumsum([len(t) for t in targets])
crit.tokens.arc_sort(True)
handles = (ctypes.c_int32 * Bt)()
_lib.check(L_.wfst_transducer_alignment_graphs(crit.tokens._h, crit.lexicon._h, flat.ctypes.data, offs.ctypes.data, Bt, handles))
aligns = [G.Graph(_handle=h) for h in handles]
packed = G.pack_graphs(aligns, x.device)
print("cfg4 alignment graphs: max nodes %d, max arcs %d" % (packed.max_nodes, packed.max_arcs))
gs = torch.full((Bt,), -1.0 / Bt, device="cuda")
def tdc():
    lattice_forward_backward(x, packed, grad_scale=gs, want_grad_emissions=True, want_grad_weights=False)

for name in ("default (lean, two-block cluster)", "lean single block", "generic"):
    old = L_.wfst_debug_force_generic_lattice({"d": 0, "l": 2, "g": 1}[name[0]])
    print("%s: cfg3 ASG step %.3f ms (Function + backward), %.3f ms (wfst_asg_forward_backward at the C ABI)"
          % (name, ev_time(asg), ev_time(asg_abi, 20)), flush=True)
    print("%s: cfg4 transducer lattice kernel (B=64, T=1000, C=%d) %.3f ms" % (name, Ct, ev_time(tdc, 3)), flush=True)
    o2 = L_.wfst_debug_force_generic_ctc(1)
    print("%s: cfg2 CTC on the log-semiring kernel only %.3f ms" % (name, ev_time(ctc, 3)), flush=True)
    L_.wfst_debug_force_generic_ctc(o2)
    L_.wfst_debug_force_generic_lattice(old)
