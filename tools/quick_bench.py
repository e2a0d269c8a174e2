"""Quick device-side timing of the rollout (not the contract bench; see bench.py)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from satflow_b200 import ConvLSTM, _lib

def run(B, tin, tout, hid, HW, train, iters=3, dtype="fp16"):
    torch.manual_seed(0)
    net = ConvLSTM(12, hid, 12, operand_dtype=dtype).cuda()
    x = torch.randn(B, tin, 12, HW, HW, device="cuda")
    tgt = torch.rand(B, tout, 12, HW, HW, device="cuda")
    def step():
        if train:
            y = net(x, tout)
            loss = torch.nn.functional.mse_loss(y.permute(0, 2, 1, 3, 4), tgt)
            loss.backward()
        else:
            with torch.no_grad():
                net(x, tout)
    for _ in range(2): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = _lib.lib().clstm_launch_count()
    e0.record()
    for _ in range(iters): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    n1 = _lib.lib().clstm_launch_count()
    flops = 0
    for k in range(4):
        cin = 12 if k == 0 else hid
        T = tin if k < 2 else tout
        flops += 2 * B * HW * HW * (cin + hid) * 4 * hid * 9 * T
    flops += 2 * B * tout * HW * HW * hid * 12 * 9
    if train: flops *= 3
    print(f"B={B} T={tin}/{tout} hid={hid} {HW}x{HW} train={train} {dtype}: {ms:.2f} ms/step, "
          f"{B*(tin+tout)/ms*1e3:.0f} frames/s, {flops/ms/1e9:.1f} TFLOP/s, launches/step={(n1-n0)/iters:.0f}", flush=True)
    net.release_plans()

if __name__ == "__main__":
    run(2, 4, 4, 32, 64, False)
    run(16, 12, 24, 64, 256, False)
    run(16, 12, 24, 64, 256, False, dtype="bf16")
    run(4, 12, 24, 64, 256, True)
    run(16, 12, 24, 64, 256, True, iters=2)
