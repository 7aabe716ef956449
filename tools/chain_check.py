"""Quick check of the chain-split CTC kernel: hazards + error vs the float64 DP + time, several shapes."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from _capi import ctc_capi
import dp_numpy
from gtn_applications_b200 import _lib as _l
_l.lib().wfst_debug_force_generic_ctc(int(os.environ.get('WFST_CTC_HOOK', '5')))
def run(B, T, C, L, scale=1.0, ragged=False, nchk=4, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, T, C, generator=g) * scale
    if ragged:
        tg = [torch.randint(C - 1, (int(n),), generator=g).tolist() for n in torch.randint(0, L + 1, (B,), generator=g)]
    else:
        tg = torch.randint(C - 1, (B, L), generator=g).tolist()
    e = torch.log_softmax(x, 2)
    losses, mean, grad, flags = ctc_capi(e.cuda(), tg, C - 1)
    worst = 0.0; lrel = 0.0
    for b in range(min(nchk, B)):
        Z, gZ = dp_numpy.ctc_dense_one(e[b].numpy().astype(np.float64), tg[b], C - 1)
        want = -gZ / B
        if not np.isfinite(Z):
            continue
        lrel = max(lrel, abs(losses[b] + Z) / max(abs(Z), 1e-30))
        sc = np.abs(want).max()
        worst = max(worst, float((np.abs(grad[b] - want) / (1e-4 * np.abs(want) + 1e-4 * sc)).max()))
    bits = " ".join("%d:%d" % (k, int(((flags & k) != 0).sum())) for k in (1, 2, 4, 8, 16))
    print("B=%d T=%d C=%d L=%d s=%g ragged=%d: flagged %d/%d [%s] loss rel %.2e grad err/tol %.3f" % (
        B, T, C, L, scale, ragged, int((flags != 0).sum()), B, bits, lrel, worst), flush=True)
if __name__ == "__main__":
    for sh in [(4, 50, 12, 7), (4, 1, 5, 0), (3, 17, 9, 8), (8, 100, 30, 40), (8, 300, 30, 100), (4, 333, 40, 150)]:
        run(*sh); run(*sh, ragged=True)
    run(16, 1000, 30, 176)
    run(16, 1000, 30, 176, scale=2.0)
    run(16, 1000, 30, 176, scale=3.0)
    run(16, 1000, 30, 176, scale=4.0)
    run(8, 1500, 80, 264)
    run(8, 777, 100, 300)
