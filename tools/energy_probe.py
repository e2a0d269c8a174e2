"""Energy per launch of the hot kernels: each kernel runs back to back for ~2.5 s on real data while nvidia-smi power /
SM clock are sampled (5 Hz), next to two calibration loads on the same box in the same process: cuBLAS bf16 8192^3
(what MEASURED_PEAKS' sustained tensor peak is) and a device copy (its HBM peak).  Reports W, MHz, us / launch,
J / launch and pJ / FLOP — the evidence behind DESIGN.md's "the step is bounded by the board's power, not by a pipe".
Usage: python tools/energy_probe.py [seconds per kernel]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import ClockSampler
from satflow_b200 import ConvLSTM


def loop(fn, seconds):
    """Run fn back to back for `seconds`; returns (us per call, sampler summary)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    # calibrate the call count so that the host never waits inside the sampled window
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    n = max(20, int(seconds * 1e3 / (e0.elapsed_time(e1) / 10)))
    s = ClockSampler(0)
    time.sleep(0.3)
    s.start()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    summ = s.stop()
    return e0.elapsed_time(e1) * 1e3 / n, summ


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 2.5
    B, hid, HW, tin, tout = 16, 64, 256, 4, 6
    torch.manual_seed(0)
    net = ConvLSTM(12, hid, 12).cuda()
    x = torch.randn(B, tin, 12, HW, HW, device="cuda")
    tgt = torch.rand(B, tout, 12, HW, HW, device="cuda")
    y = net(x, tout)
    torch.nn.functional.mse_loss(y.permute(0, 2, 1, 3, 4), tgt).backward()
    plan = [p for p in net._plans.values() if p.training][0]
    torch.cuda.synchronize()
    fl = 2 * B * HW * HW * (hid + hid) * 4 * hid * 9
    rows = []

    def report(name, us, summ, flops=None, nbytes=None):
        w = summ.get("power_w_median") or float("nan")
        r = {"kernel": name, "us": round(us, 1), "power_w": w, "sm_mhz": summ.get("sm_mhz"),
             "reasons": summ.get("reasons"), "joule_per_launch": round(w * us * 1e-6, 4), "samples": summ.get("samples")}
        if flops:
            r["tflops"] = round(flops / us / 1e6, 1)
            r["pj_per_flop"] = round(w * us * 1e-6 / flops * 1e12, 3)
        if nbytes:
            r["gbs"] = round(nbytes / us / 1e3, 1)
            r["pj_per_byte"] = round(w * us * 1e-6 / nbytes * 1e12, 1)
        rows.append(r)
        print(json.dumps(r), flush=True)

    # idle floor
    s = ClockSampler(0)
    s.start()
    time.sleep(1.5)
    idle = s.stop()
    print(json.dumps({"kernel": "idle", "power_w": idle.get("power_w_median"), "sm_mhz": idle.get("sm_mhz")}), flush=True)

    a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
    b = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
    us, summ = loop(lambda: torch.matmul(a, b), seconds)
    report("cuBLAS bf16 8192^3 (random normal operands)", us, summ, flops=2 * 8192 ** 3)
    ah, bh = a.half(), b.half()
    us, summ = loop(lambda: torch.matmul(ah, bh), seconds)
    report("cuBLAS fp16 8192^3 (random normal operands)", us, summ, flops=2 * 8192 ** 3)
    del a, b, ah, bh
    src = torch.empty(1 << 30, device="cuda", dtype=torch.bfloat16)
    dst = torch.empty_like(src)
    us, summ = loop(lambda: dst.copy_(src), seconds)
    report("device copy 2 GiB -> 2 GiB", us, summ, nbytes=2 * src.numel() * 2)
    del src, dst

    px = B * HW * HW * hid
    # algorithmic bytes per launch with the default storage formats (16-bit c, dc, dh; c' recomputed): DESIGN.md §4
    for kind, name, nbytes in (("cell_fwd", "fused cell step (training variant)", px * 18),
                               ("dgrad_fused", "dgradT + fused gate gradient", px * 34),
                               ("dgrad", "dgradT alone", px * 16), ("wgrad", "wgrad (halo rows)", px * 12),
                               ("gate_grad", "gate gradient alone", px * 40)):
        try:
            us, summ = loop(lambda: plan.profile_kernel(kind, 3, 2), seconds)
        except Exception as e:  # a kind the current schedule does not use
            print(json.dumps({"kernel": name, "error": str(e)[:120]}), flush=True)
            continue
        report(name, us, summ, flops=None if kind == "gate_grad" else fl, nbytes=nbytes)
    out = os.path.join(ROOT, "gpurun_out", os.environ.get("PROBE_OUT", "energy_probe.json"))
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(out, "w") as f:
        json.dump({"idle": idle, "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
