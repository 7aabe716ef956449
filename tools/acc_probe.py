"""Accuracy probe (GPU box): CUDA CTC / ASG vs the float64 numpy DP oracle at large T."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import dp_numpy
from gtn_applications_b200.criterions.ctc import CTCLoss
from gtn_applications_b200.criterions.asg import ASGLoss

def report(name, g, w, extra=""):
    scale = np.abs(w).max()
    rel = np.abs(g - w) / np.maximum(np.abs(w), 1e-3 * scale)
    print(f"{name}: max abs {np.abs(g-w).max():.2e} (scale {scale:.2e}) max rel(floored 1e-3) {rel.max():.2e} {extra}")

import ctypes, time
from gtn_applications_b200 import _lib, _runtime as rt

def hazards(B, T, C, L):
    ws = rt._workspaces[torch.device("cuda", 0)]
    flags = (ctypes.c_int32 * B)()
    _lib.check(_lib.lib().wfst_debug_ctc_hazards(ws.data_ptr(), B, T, C, L, flags))
    return sum(1 for f in flags if f > 0), list(flags)[:4]

torch.manual_seed(0)
for (B, T, C, L, lsm) in [(2, 5, 4, 2, True), (3, 40, 6, 7, True), (8, 257, 31, 60, True), (8, 1000, 30, 176, True), (8, 1000, 30, 176, False), (4, 1500, 80, 264, True)]:
    x = torch.randn(B, T, C)
    lp = torch.log_softmax(x, 2) if lsm else x
    tg = torch.randint(C - 2, (B, L)).tolist()
    a = lp.clone().cuda().requires_grad_(True)
    loss = CTCLoss(a, tg, C - 1, "none"); loss.backward()
    ref = dp_numpy.ctc(lp.numpy(), tg, C - 1, "none")
    g = a.grad.cpu().numpy().astype(np.float64)
    report(f"CTC B{B} T{T} C{C} L{L} lsm={lsm} loss rel {abs(loss.item()-ref['loss'])/abs(ref['loss']):.1e}", g, ref["grad"],
           f"rowsum dev {np.abs(g.sum(2)*B+1).max():.1e} hazards {hazards(B, T, C, L)}")
for (B, T, C, L) in [(3, 64, 80, 20), (2, 120, 30, 50), (4, 1000, 30, 176)]:
    e = torch.randn(B, T, C); tr = torch.randn(C + 1, C)
    tg = torch.randint(C, (B, L)).tolist()
    a = e.clone().cuda().requires_grad_(True); t = tr.clone().cuda().requires_grad_(True)
    loss = ASGLoss(a, t, tg, "none"); loss.backward()
    ref = dp_numpy.asg(e.numpy(), tr.numpy(), tg, "none")
    report(f"ASG B{B} T{T} C{C} L{L} loss rel {abs(loss.item()-ref['loss'])/abs(ref['loss']):.1e} gradE", a.grad.cpu().numpy().astype(np.float64), ref["grad"])
    report("      gradTrans", t.grad.cpu().numpy().astype(np.float64), ref["grad_transitions"])
