"""Short target for ncu: one training step (fwd + bwd) of the bench workload, optionally preceded by warm-up."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from satflow_b200 import ConvLSTM

B = int(os.environ.get("PROF_B", "16"))
steps = int(os.environ.get("PROF_STEPS", "1"))
torch.manual_seed(0)
net = ConvLSTM(12, 64, 12).cuda()
if os.environ.get("PROF_SMALL") == "cfg0":  # BASELINE configs[0] on the 2+2 architecture
    steps = 3
    net = ConvLSTM(12, 32, 12).cuda()
    x = torch.randn(2, 4, 12, 64, 64, device="cuda")
    tgt = torch.rand(2, 4, 12, 64, 64, device="cuda")
elif os.environ.get("PROF_SMALL"):  # a shape that takes the persistent forward chain (128 tiles <= 148 SMs)
    steps = 3
    x = torch.randn(1, 12, 12, 128, 128, device="cuda")
    tgt = torch.rand(1, 24, 12, 128, 128, device="cuda")
else:
    x = torch.randn(B, 12, 12, 256, 256, device="cuda")
    tgt = torch.rand(B, 24, 12, 256, 256, device="cuda")
for _ in range(steps):
    y = net(x, tgt.shape[1])
    loss = torch.nn.functional.mse_loss(y.permute(0, 2, 1, 3, 4), tgt)
    loss.backward()
torch.cuda.synchronize()
print("loss", loss.item())
