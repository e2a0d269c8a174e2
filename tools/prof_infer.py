"""Short target for ncu: inference rollouts of the bench workload (PROF_STEPS, default 1)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from satflow_b200 import ConvLSTM

B = int(os.environ.get("PROF_B", "16"))
torch.manual_seed(0)
net = ConvLSTM(12, 64, 12).cuda()
x = torch.randn(B, 12, 12, 256, 256, device="cuda")
with torch.no_grad():
    for _ in range(int(os.environ.get("PROF_STEPS", "1"))):
        y = net(x, 24)
torch.cuda.synchronize()
print("mean", y.mean().item())
