import os, sys, torch
sys.path.insert(0, "/root/repo")
from gtn_applications_b200.criterions.ctc import CTCLogitsLoss
torch.manual_seed(0)
B, T, C, L = 256, 1000, 30, 176
x = torch.randn(B, T, C, device="cuda").requires_grad_(True)
tg = torch.randint(C - 2, (B, L)).tolist()
for i in range(2):
    loss = CTCLogitsLoss(x, tg, C - 1, "none"); loss.backward(); torch.cuda.synchronize()
