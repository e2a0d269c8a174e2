"""Times a full inference rollout and one training step; used for A/B of knobs that affect the small kernels."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from satflow_b200 import EncoderDecoderConvLSTM
torch.manual_seed(0)
m = EncoderDecoderConvLSTM(hidden_dim=64, input_channels=12, out_channels=12, forecast_steps=24).cuda()
x = torch.randn(16, 12, 12, 256, 256, device="cuda"); y = torch.rand(16, 24, 12, 256, 256, device="cuda")
def t(fn, n=3):
    for _ in range(2): fn()
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
def inf():
    with torch.no_grad(): m(x, 24)
def trn():
    m.zero_grad(set_to_none=True); m.training_step((x, y), 0).backward()
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("CLSTM_"))
print(f"[{tag}] infer {t(inf):.2f} ms  train {t(trn):.2f} ms", flush=True)
