import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gtn_applications_b200.criterions.ctc import CTCLoss
from gtn_applications_b200 import _lib
_lib.lib().wfst_debug_force_generic_ctc(int(os.environ.get('WFST_CTC_HOOK','0')))
torch.manual_seed(0)
B, T, C, L = int(os.environ.get('B','256')), 1000, 30, 176
lp = torch.log_softmax(torch.randn(B, T, C, device="cuda"), 2).requires_grad_(True)
tg = torch.randint(C - 2, (B, L)).tolist()
for i in range(2):
    loss = CTCLoss(lp, tg, C - 1, "none"); loss.backward(); torch.cuda.synchronize()
