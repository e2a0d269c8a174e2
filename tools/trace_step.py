"""In-situ per-kernel timing of the bench workload's training step (and an inference rollout) through the library's
launch trace (clstm_trace_enable / clstm_trace_report): CUDA events after every launch, under the real clocks of the
step.  Usage: python tools/trace_step.py [steps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from satflow_b200 import EncoderDecoderConvLSTM, _lib

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
torch.manual_seed(0)
m = EncoderDecoderConvLSTM(hidden_dim=64, input_channels=12, out_channels=12, forecast_steps=24).cuda()
x = torch.randn(16, 12, 12, 256, 256, device="cuda")
y = torch.rand(16, 24, 12, 256, 256, device="cuda")


def train():
    m.zero_grad(set_to_none=True)
    m.training_step((x, y), 0).backward()


for _ in range(2):
    train()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(steps):
    train()
b.record()
torch.cuda.synchronize()
print(f"untraced: {a.elapsed_time(b) / steps:.2f} ms per training step")
_lib.trace_enable(8192)
a.record()
for _ in range(steps):
    train()
b.record()
torch.cuda.synchronize()
print(f"traced:   {a.elapsed_time(b) / steps:.2f} ms per training step ({steps} steps in the table below)")
print(_lib.trace_report())
with torch.no_grad():
    m(x, 24)
    torch.cuda.synchronize()
    _lib.trace_report()
    m(x, 24)
print("inference rollout:")
print(_lib.trace_report())
_lib.trace_enable(0)
