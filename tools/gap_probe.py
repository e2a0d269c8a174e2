"""GPU busy time vs wall time of one training step (torch.profiler / CUPTI sees the ctypes-launched kernels too)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
from satflow_b200 import EncoderDecoderConvLSTM

torch.manual_seed(0)
m = EncoderDecoderConvLSTM(hidden_dim=64, input_channels=12, out_channels=12, forecast_steps=24).cuda()
opt = torch.optim.Adam(m.parameters(), lr=1e-4, fused=True)
x = torch.randn(16, 12, 12, 256, 256, device="cuda")
y = torch.rand(16, 24, 12, 256, 256, device="cuda")
def step():
    opt.zero_grad(set_to_none=False)
    loss = m.training_step((x, y), 0)
    loss.backward()
    opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
busy = sum(e.time_range.end - e.time_range.start for e in evs)
span = evs[-1].time_range.end - evs[0].time_range.start
gaps = []
for a, b in zip(evs, evs[1:]):
    g = b.time_range.start - a.time_range.end
    if g > 0: gaps.append((g, a.name[:40], b.name[:40]))
print(f"kernels {len(evs)}  span {span/1e3:.2f} ms  busy {busy/1e3:.2f} ms  idle {100*(1-busy/span):.1f}%")
gaps.sort(reverse=True)
for g, a, b in gaps[:12]: print(f"  gap {g:8.1f} us after {a} before {b}")
import collections
agg = collections.defaultdict(float)
for e in evs: agg[e.name[:50]] += e.time_range.end - e.time_range.start
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]: print(f"  {v/2e3:8.2f} ms/step  {k}")
