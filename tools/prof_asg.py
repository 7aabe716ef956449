import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gtn_applications_b200.criterions.asg import ASGLoss
torch.manual_seed(0)
B, T, C, L = 256, 1000, 30, 176
e = torch.randn(B, T, C, device="cuda", requires_grad=True)
tr = torch.randn(C + 1, C, device="cuda", requires_grad=True)
tg = torch.randint(C, (B, L)).tolist()
for i in range(2):
    e.grad = None; tr.grad = None
    ASGLoss(e, tr, tg, "mean").backward(); torch.cuda.synchronize()
