"""GPU check of the CTC path: errors against the float64 numpy DP on several shapes,
scaled-kernel fallback flags, and event timing at cfg2.  Usage: python tools/check_pair.py [kind]
(kind: 0 default dispatch, 2 = no paired kernel, 1 = log-semiring only)."""
import os, sys, time, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import dp_numpy
from gtn_applications_b200 import _lib, _runtime as rt
from gtn_applications_b200.criterions.ctc import CTCLoss
kind = int(sys.argv[1]) if len(sys.argv) > 1 else 0
_lib.lib().wfst_debug_force_generic_ctc(kind)

def hazards(B, T, C, L):
    flags = (np.zeros(B, dtype=np.int32))
    ws = rt.workspace(torch.device("cuda:0"), _lib.lib().wfst_ctc_workspace_bytes(B, T, C, L))
    _lib.check(_lib.lib().wfst_debug_ctc_hazards(ws.data_ptr(), B, T, C, L, flags.ctypes.data))
    return flags

def one(B, T, C, lens, lsm=True, seed=0, blank=None):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((B, T, C)).astype(np.float32) * 2
    if lsm:
        x = x - np.log(np.exp(x.astype(np.float64)).sum(2, keepdims=True))
        x = x.astype(np.float32)
    blank = C - 1 if blank is None else blank
    labs = [c for c in range(C) if c != blank]
    tg = [rng.choice(labs, size=n).tolist() for n in lens]
    a = torch.tensor(x, device="cuda").requires_grad_(True)
    loss = CTCLoss(a, tg, blank, "none"); loss.backward(); torch.cuda.synchronize()
    hz = hazards(B, T, C, max(lens) if lens else 0)
    ref = dp_numpy.ctc(x, tg, blank, "none")
    g = a.grad.cpu().numpy()
    gerr = np.abs(g - ref["grad"]).max() / max(np.abs(ref["grad"]).max(), 1e-30)
    lerr = abs(loss.item() - ref["loss"]) / max(abs(ref["loss"]), 1e-30) if np.isfinite(ref["loss"]) else (0.0 if not np.isfinite(loss.item()) else 1.0)
    ok = gerr < 1e-4 and lerr < 1e-4
    print(f"B={B} T={T} C={C} lens={lens[:6]} lsm={lsm}: loss relerr {lerr:.2e} grad err/max {gerr:.2e} hazards {hz.tolist()[:8]} {'ok' if ok else 'FAIL'}", flush=True)
    return ok

ok = True
ok &= one(1, 40, 6, [7])
ok &= one(2, 40, 6, [7, 5])
ok &= one(3, 12, 6, [4, 0, 6])
ok &= one(5, 33, 9, [1, 16, 7, 0, 11])
ok &= one(4, 150, 28, [20] * 4, lsm=False)
ok &= one(2, 64, 5, [30, 31], lsm=False)
ok &= one(6, 257, 31, [40, 3, 77, 128, 0, 64])
ok &= one(3, 100, 30, [191, 100, 63])
ok &= one(2, 300, 50, [255, 130])
ok &= one(2, 200, 100, [60, 61], blank=0)
ok &= one(4, 1000, 30, [176] * 4)
ok &= one(4, 1003, 30, [176, 150, 100, 10], lsm=False)
print("ALL OK" if ok else "SOME FAILED")

torch.manual_seed(0)
B, T, C, L = 256, 1000, 30, 176
lp = torch.log_softmax(torch.randn(B, T, C, device="cuda"), 2).requires_grad_(True)
tg = torch.randint(C - 2, (B, L)).tolist()
for i in range(3):
    loss = CTCLoss(lp, tg, C - 1, "none"); loss.backward()
torch.cuda.synchronize()
hz = hazards(B, T, C, L)
print("cfg2 hazards nonzero:", int((hz != 0).sum()), "loss", loss.item())
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for i in range(20):
    lp.grad = None
    loss = CTCLoss(lp, tg, C - 1, "none"); loss.backward()
ev[1].record(); torch.cuda.synchronize()
print("cfg2 ms/iter (python path, 20 iters): %.4f" % (ev[0].elapsed_time(ev[1]) / 20))
