"""In-situ launch trace of a small rollout (BASELINE configs[0] on the 2+2 architecture): where a launch-bound step goes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from satflow_b200 import ConvLSTM, _lib

B, tin, tout, hid, hw = (int(a) for a in (sys.argv[1:6] if len(sys.argv) > 5 else (2, 4, 4, 32, 64)))
torch.manual_seed(0)
net = ConvLSTM(12, hid, 12).cuda()
x = torch.randn(B, tin, 12, hw, hw, device="cuda")
tgt = torch.rand(B, tout, 12, hw, hw, device="cuda")

def train():
    net.zero_grad(set_to_none=True)
    torch.nn.functional.mse_loss(net(x, tout).permute(0, 2, 1, 3, 4), tgt).backward()

for _ in range(5):
    train()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    train()
b.record()
torch.cuda.synchronize()
print(f"B={B} {tin}/{tout} hid {hid} {hw}x{hw}: {a.elapsed_time(b) / 20:.3f} ms per training step (untraced)")
import time
t0 = time.perf_counter()
for _ in range(20):
    train()
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"host enqueue time per step: {(t1 - t0) / 20 * 1e3:.3f} ms")
_lib.trace_enable(8192)
for _ in range(4):
    train()
torch.cuda.synchronize()
print(_lib.trace_report())
_lib.trace_enable(0)
