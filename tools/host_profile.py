"""cProfile of the host side of a small (launch-bound) training step: where the Python + ctypes time goes."""
import cProfile, os, pstats, sys, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from satflow_b200 import ConvLSTM

torch.manual_seed(0)
net = ConvLSTM(12, 32, 12).cuda()
x = torch.randn(2, 4, 12, 64, 64, device="cuda")
tgt = torch.rand(2, 4, 12, 64, 64, device="cuda")

def train():
    net.zero_grad(set_to_none=True)
    torch.nn.functional.mse_loss(net(x, 4).permute(0, 2, 1, 3, 4), tgt).backward()

for _ in range(20):
    train()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(300):
    train()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:9000])
