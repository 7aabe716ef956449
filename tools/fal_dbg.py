import os, sys, torch
sys.path.insert(0, os.getcwd())
from gtn_applications_b200.criterions.asg import ASGLoss
torch.manual_seed(0)
B, T, C, L = int(os.environ.get("B","4")), int(os.environ.get("T","100")), int(os.environ.get("C","10")), int(os.environ.get("L","12"))
e = torch.randn(B, T, C, device="cuda", requires_grad=True)
tr = torch.randn(C + 1, C, device="cuda", requires_grad=True)
tg = torch.randint(C, (B, L)).tolist()
loss = ASGLoss(e, tr, tg, "mean"); loss.backward(); torch.cuda.synchronize()
print("loss", loss.item())
