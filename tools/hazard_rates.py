"""Share of utterances the scaled CTC kernel hands to the log-semiring kernel, by reason bit,
as the emissions get steeper (log_softmax(scale * randn)) and for an aligned 'trained-like' case."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gtn_applications_b200 import _lib, _runtime as rt
from gtn_applications_b200.criterions.ctc import CTCLoss

def hazards(B, T, C, L):
    flags = np.zeros(B, dtype=np.int32)
    ws = rt.workspace(torch.device("cuda:0"), _lib.lib().wfst_ctc_workspace_bytes(B, T, C, L))
    _lib.check(_lib.lib().wfst_debug_ctc_hazards(ws.data_ptr(), B, T, C, L, flags.ctypes.data))
    return flags

B, T, C, L = int(os.environ.get("WFST_B", "64")), 1000, 30, 176
_lib.lib().wfst_debug_force_generic_ctc(int(os.environ.get("WFST_CTC_HOOK", "0")))
if os.environ.get("WFST_CHAIN_CFG"):
    _lib.lib().wfst_debug_ctc_chain_config(*[int(v) for v in os.environ["WFST_CHAIN_CFG"].split(",")])
torch.manual_seed(0)
tg = torch.randint(C - 2, (B, L))
for scale in (1, 2, 3, 4, 6, 8, 12):
    lp = torch.log_softmax(torch.randn(B, T, C, device="cuda") * scale, 2).requires_grad_(True)
    CTCLoss(lp, tg.tolist(), C - 1, "none").backward(); torch.cuda.synchronize()
    hz = hazards(B, T, C, L)
    print("scale %2d: flagged %3d/%d  bits: " % (scale, int((hz != 0).sum()), B) +
          " ".join("%d:%d" % (b, int(((hz & b) != 0).sum())) for b in (1, 2, 4, 8, 16)), flush=True)
# trained-like: a forced alignment gets +margin on its label
for margin in (3, 6, 10, 15):
    x = torch.randn(B, T, C)
    for b in range(B):
        pos = np.sort(np.random.default_rng(b).choice(T, size=L, replace=False))
        x[b, :, C - 1] += margin          # blank dominates ...
        x[b, pos, C - 1] -= margin
        x[b, pos, tg[b]] += margin        # ... except at L spike frames
    lp = torch.log_softmax(x.cuda(), 2).requires_grad_(True)
    loss = CTCLoss(lp, tg.tolist(), C - 1, "none"); loss.backward(); torch.cuda.synchronize()
    hz = hazards(B, T, C, L)
    print("spiky margin %2d: loss %.1f flagged %3d/%d  bits: " % (margin, loss.item(), int((hz != 0).sum()), B) +
          " ".join("%d:%d" % (b, int(((hz & b) != 0).sum())) for b in (1, 2, 4, 8, 16)), flush=True)
