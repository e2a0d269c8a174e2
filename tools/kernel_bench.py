"""Per-kernel device timing of the bench workload's cell kernels through the C-ABI measurement hook
(clstm_plan_profile_kernel).  Usage: python tools/kernel_bench.py [B] ; knobs via CLSTM_* env vars."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from satflow_b200 import ConvLSTM

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    kinds = sys.argv[2].split(",") if len(sys.argv) > 2 else ["cell_fwd", "gate_grad", "dgrad", "wgrad"]
    hid, HW, tin, tout = 64, 256, 4, 6
    torch.manual_seed(0)
    net = ConvLSTM(12, hid, 12).cuda()
    x = torch.randn(B, tin, 12, HW, HW, device="cuda")
    tgt = torch.rand(B, tout, 12, HW, HW, device="cuda")
    y = net(x, tout)
    torch.nn.functional.mse_loss(y.permute(0, 2, 1, 3, 4), tgt).backward()
    plan = [p for p in net._plans.values() if p.training][0]
    # Schedule knobs are read when a plan is created (the rollout above), so CLSTM_* set in the environment of this
    # process apply to both the rollout and the timed launches; the states hold real data (operand values change the
    # power draw and therefore the clocks of a power-capped B200).
    torch.cuda.synchronize()
    fl = 2 * B * HW * HW * (hid + hid) * 4 * hid * 9
    tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("CLSTM_"))
    for kind in kinds:
        for cell in ((0, 3) if kind in ("cell_fwd", "wgrad") else (3,)):
            for _ in range(3):
                plan.profile_kernel(kind, cell, int(os.environ.get("KB_STEP", "2")) if cell == 3 else 2)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 20
            e0.record()
            for _ in range(reps):
                plan.profile_kernel(kind, cell, int(os.environ.get("KB_STEP", "2")) if cell == 3 else 2)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            f = fl if cell == 3 else 2 * B * HW * HW * (12 + hid) * 4 * hid * 9
            extra = f"{f / ms / 1e9:7.1f} TFLOP/s" if kind != "gate_grad" else ""
            print(f"[{tag}] {kind:9s} cell {cell}: {ms * 1e3:8.1f} us  {extra}", flush=True)

if __name__ == "__main__":
    main()
