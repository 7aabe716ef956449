import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import dp_numpy
from gtn_applications_b200.criterions.ctc import CTCLoss
torch.manual_seed(0)
B, T, C, L = 1, 40, 6, 7
if len(sys.argv) > 1: B, T, C, L = map(int, sys.argv[1:5])
x = torch.randn(B, T, C); lp = torch.log_softmax(x, 2)
tg = torch.randint(C - 2, (B, L)).tolist()
a = lp.clone().cuda().requires_grad_(True)
loss = CTCLoss(a, tg, C - 1, "none"); loss.backward(); torch.cuda.synchronize()
ref = dp_numpy.ctc(lp.numpy(), tg, C - 1, "none")
print("loss", loss.item(), ref["loss"], "grad err", np.abs(a.grad.cpu().numpy() - ref["grad"]).max())
