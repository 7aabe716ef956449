"""cProfile of the Python side of one CTC step through the Function (cfg2, tensor targets, inputs
already on the device): what stands between bench.py's e2e number and the PCIe ceiling."""
import cProfile, os, pstats, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gtn_applications_b200.criterions.ctc import CTCLoss
torch.manual_seed(0)
B, T, C, L = 256, 1000, 30, 176
lp = torch.log_softmax(torch.randn(B, T, C, device="cuda"), 2)
tg = torch.randint(C - 2, (B, L))
def step():
    x = lp.detach().requires_grad_(True)
    loss = CTCLoss(x, tg, C - 1, "none")
    loss.backward()
    return loss
for _ in range(5): step()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(100): step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
