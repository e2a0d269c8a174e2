#!/usr/bin/env python
"""Benchmark of the ConvLSTM encoder-forecaster hot path (BASELINE.json metric: rollout frames/s, fwd+bwd).

  python bench.py --gpus N --steps K --warmup W              our arm (one process per GPU; torchrun for N>1)
  python bench.py --impl reference --gpus N --steps K ...    the reference's CPU implementation of the same path

One step = zero_grad + forward rollout + MSE loss + backward (BPTT) + gradient all-reduce (N>1) + Adam step
on one batch of synthetic 12-channel sequences, hid 64, 256x256, 12 in / 24 out, batch 16 per GPU (weak scaling;
at N=8 this is BASELINE configs[2] exactly: global batch 128).  frames/s = B * (T_in + T_out) / step time.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "ConvLSTM rollout frames/s (fwd+bwd)"
CFG = dict(hidden=64, channels=12, out_channels=12, hw=256, t_in=12, t_out=24, batch_per_gpu=16)


def algorithmic_flops_fwd(B, t_in, t_out, cin, hid, cout, H, W, k=3, n_layers=2):
    """2*B*H*W*(Cin_x+hid)*4hid*k*k per cell step (unpadded), head 2*B*T_out*H*W*hid*C_out*9."""
    f = 0
    for c in range(2 * n_layers):
        cx = cin if c == 0 else hid
        T = t_in if c < n_layers else t_out
        f += 2 * B * H * W * (cx + hid) * 4 * hid * k * k * T
    return f + 2 * B * t_out * H * W * hid * cout * 9


def cell_step_flops(B, cx, hid, H, W, k=3):
    return 2 * B * H * W * (cx + hid) * 4 * hid * k * k


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            d = json.load(f)
        return dict(bf16_burst=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    hbm=d["hbm_gbs"], source="measured (MEASURED_PEAKS.json)")
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_sample(steps: int, warmup: int, threads: int):
    """The reference's CPU fp32 path (torch autograd over cat/Conv2d/sigmoid/tanh, exactly the op sequence of
    layers/ConvLSTM.py:42-57 and conv_lstm.py:171-203, restated in oracle/convlstm_oracle.py because
    /root/reference does not exist on the GPU box).  Bounded sample of the bench workload: same hid/channels/
    256x256, batch 1, 4 in / 8 out steps; frames/s scales linearly in batch and steps."""
    from oracle import convlstm_oracle as O

    torch.set_num_threads(threads)
    B, t_in, t_out = 1, 4, 8
    g = torch.Generator().manual_seed(1234)
    p = {k: v.requires_grad_(True) for k, v in O.init_params(CFG["channels"], CFG["hidden"], CFG["out_channels"], seed=0).items()}
    x = torch.randn(B, t_in, CFG["channels"], CFG["hw"], CFG["hw"], generator=g)
    tgt = torch.rand(B, t_out, CFG["out_channels"], CFG["hw"], CFG["hw"], generator=g)
    opt = torch.optim.Adam(list(p.values()), lr=1e-3)

    def step():
        opt.zero_grad(set_to_none=True)
        y, _ = O.rollout_forward(x, p, t_out)
        loss = torch.nn.functional.mse_loss(y.permute(0, 2, 1, 3, 4), tgt)
        loss.backward()
        opt.step()
        return loss.item()

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return dict(value=B * (t_in + t_out) / dt, ms_per_step=dt * 1e3,
                sample=f"B=1, {t_in} in / {t_out} out, hid {CFG['hidden']}, {CFG['hw']}x{CFG['hw']}, fwd+bwd+Adam, "
                       f"{steps} timed steps after {warmup} warm-up (torch {torch.__version__} CPU fp32)")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    r = cpu_reference_sample(max(1, args.steps), max(1, min(args.warmup, 1)), threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": r["value"], "unit": "frames/s", "cores": threads, "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(n):
    return {
        "workload": (f"encoder-forecaster ConvLSTM (2 enc + 2 dec cells, 3x3) hid {CFG['hidden']}, {CFG['channels']}ch "
                     f"{CFG['hw']}x{CFG['hw']}, {CFG['t_in']} in / {CFG['t_out']} out, batch {CFG['batch_per_gpu']}/GPU, "
                     "fwd+bwd+grad-allreduce+Adam (BASELINE configs[2]; shape of configs[1])"),
        "global_batch": CFG["batch_per_gpu"] * n, "parallelism": f"dp{n}",
        "l2": "working set (71 GB of saved states per step) >> 126 MB L2; no flush needed",
    }


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch.distributed as dist

    from satflow_b200 import EncoderDecoderConvLSTM, _lib
    from satflow_b200.distributed import FlatGradBucket, broadcast_parameters

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun (--nproc-per-node {args.gpus})")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    _lib.check(L.clstm_device_check(local))

    B, t_in, t_out = CFG["batch_per_gpu"], CFG["t_in"], CFG["t_out"]
    C, Co, hid, HW = CFG["channels"], CFG["out_channels"], CFG["hidden"], CFG["hw"]
    torch.manual_seed(0)
    model = EncoderDecoderConvLSTM(hidden_dim=hid, input_channels=C, out_channels=Co, forecast_steps=t_out, lr=1e-4)
    model.model.operand_dtype = args.dtype
    model = model.to(dev)
    broadcast_parameters(model)
    bucket = FlatGradBucket(model.parameters())
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)
    g = torch.Generator().manual_seed(1234 + rank)
    x_host = torch.randn(B, t_in, C, HW, HW, generator=g).pin_memory()
    y_host = torch.rand(B, t_out, Co, HW, HW, generator=g).pin_memory()
    x_dev, y_dev = x_host.to(dev), y_host.to(dev)

    def train_step(x, tgt):
        bucket.zero_()
        loss = model.training_step((x, tgt), 0)  # conv_lstm.py:53-70: forward, MSE loss, per-frame losses
        loss.backward()
        bucket.all_reduce_mean()
        opt.step()
        return loss

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """K steps bracketed by barrier + synchronize; device time via CUDA events; max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = L.clstm_launch_count()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms / steps, (L.clstm_launch_count() - n0)

    for _ in range(max(3, args.warmup)):
        train_step(x_dev, y_dev)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_step, launches = timed(lambda: train_step(x_dev, y_dev), args.steps)
    clocks = sampler.stop() if rank == 0 else None

    # end to end through the public module API with HOST buffers: every step's x and target are copied from pinned
    # host memory inside the timed region (double-buffered on a side stream so the copy of step i+1 overlaps
    # step i, as a DataLoader with pin_memory does) and the loss is read back to the host every step.
    from satflow_b200.prefetch import DevicePrefetcher

    def e2e_run(steps):
        pf = DevicePrefetcher(((x_host, y_host) for _ in range(steps)), dev)
        out = 0.0
        for xd, td in pf:
            loss = train_step(xd, td)
            pf.done_with_current()
            out = loss.item()
        return out

    e2e_run(2)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_run(args.steps)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = t.item()
    ms_e2e /= args.steps

    # inference (BASELINE configs[1]): forward rollout only, same shapes
    extra = {}
    if rank == 0 or world > 1:
        def infer():
            with torch.no_grad():
                model(x_dev, t_out)
        for _ in range(3):
            infer()
        ms_inf, _ = timed(infer, args.steps)
        extra["inference"] = {"frames_per_s": world * B * (t_in + t_out) / ms_inf * 1e3, "ms_per_step": ms_inf,
                              "tflops": algorithmic_flops_fwd(B, t_in, t_out, C, hid, Co, HW, HW) / ms_inf / 1e9}

    # roofline of the dominant kernel (the fused cell step or the fused dgrad + gate-gradient launch, whichever takes
    # the larger share of the step), timed alone with CUDA events on the launching stream through the C-ABI
    # measurement hook; the other kernels of the step alongside.
    peaks = measured_peaks()
    roof = None
    cpu = None
    if rank == 0:
        plan = [p for p in model.model._plans.values() if p.training][0]

        def time_kernel(kind, cell, step, reps=20):
            for _ in range(3):
                plan.profile_kernel(kind, cell, step)
            torch.cuda.synchronize()
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                plan.profile_kernel(kind, cell, step)
            b_.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b_) / reps

        fl = cell_step_flops(B, hid, hid, HW, HW)  # decoder_2: K = (64 + 64) * 9
        k_ms = time_kernel("cell_fwd", 3, 5)
        ach = fl / k_ms / 1e9
        roof = {"kernel": "convgemm_kernel<EPI_LSTM> (fused cell step, training variant: also writes gates)",
                "bound": "tensor", "achieved": ach, "peak": peaks["bf16_burst"], "unit": "TFLOP/s",
                "frac": ach / peaks["bf16_burst"], "traffic": 1.43e9, "launch_ms": k_ms,
                "flops_per_launch": fl, "peak_source": peaks["source"] + " burst, kernel timed alone",
                "traffic_source": "ncu --set full dram__bytes_read+write per launch (profiles/r1_v2_ncu_full_summary.csv)"}
        others = []
        for kind, name in (("dgrad", "dgradT_kernel (data gradient)"), ("wgrad", "wgrad_kernel (weight gradient)")):
            ms = time_kernel(kind, 3, 5)
            others.append({"kernel": name, "bound": "tensor", "achieved": fl / ms / 1e9, "peak": peaks["bf16_burst"],
                           "unit": "TFLOP/s", "frac": fl / ms / 1e9 / peaks["bf16_burst"], "launch_ms": ms})
        npix = B * HW * HW
        gg_bytes = npix * hid * (4 * 2 + 4 + 4 + 2 * 4 + 2 * 4 + 4 * 2)  # gates, c_prev, c_next, 2 dh, dc r/w, dz
        ms = time_kernel("gate_grad", 3, 5)
        others.append({"kernel": "gate_grad_kernel (pointwise gate gradient)", "bound": "hbm",
                       "achieved": gg_bytes / ms / 1e6, "peak": peaks["hbm"], "unit": "GB/s",
                       "frac": gg_bytes / ms / 1e6 / peaks["hbm"], "launch_ms": ms})
        # default backward schedule: the gate gradient of the next chain step runs in the dgrad epilogue, so the
        # launch does both the GEMM flops and the pointwise pass's bytes; reported against the tensor peak
        ms = time_kernel("dgrad_fused", 3, 5)
        fused = {"kernel": "dgradT_fused_kernel (data gradient + fused gate gradient of the cell below)",
                 "bound": "tensor", "achieved": fl / ms / 1e9, "peak": peaks["bf16_burst"], "unit": "TFLOP/s",
                 "frac": fl / ms / 1e9 / peaks["bf16_burst"], "traffic": 3.31e9, "launch_ms": ms,
                 "flops_per_launch": fl, "fused_hbm_GBps": gg_bytes / ms / 1e6,
                 "peak_source": peaks["source"] + " burst, kernel timed alone",
                 "traffic_source": "ncu --set full dram__bytes_read+write per launch "
                                   "(profiles/r1_v2_ncu_full_summary.csv); algorithmic 3.24 GB"}
        # `roofline` is the kernel with the largest share of the step: 71 fused dgrad launches vs 72 cell steps
        roof["launches_per_step"] = 2 * (t_in + t_out)
        fused["launches_per_step"] = 2 * (t_in + t_out) - 1
        if fused["launch_ms"] * fused["launches_per_step"] > roof["launch_ms"] * roof["launches_per_step"]:
            roof, fused = fused, roof
        others.append(fused)
        extra["kernels"] = others
        flops_step = 3 * algorithmic_flops_fwd(B, t_in, t_out, C, hid, Co, HW, HW)
        extra["step_tflops"] = flops_step * world / ms_step / 1e9
        extra["step_frac_of_sustained_peak"] = flops_step / ms_step / 1e9 / peaks["bf16_sustained"]
        threads = os.cpu_count() or 1
        if not args.no_cpu:
            r = cpu_reference_sample(1, 1, threads)
            cpu = {"value": r["value"], "unit": "frames/s", "cores": threads, "kind": "port", "sample": r["sample"]}

    if rank == 0:
        frames = world * B * (t_in + t_out)
        line = {
            "metric": METRIC, "value": frames / ms_step * 1e3, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16" if args.dtype == "fp16" else "bf16", "data": "synthetic",
            "config": workload_config(world),
            "e2e": {"value": frames / ms_e2e * 1e3, "unit": "frames/s",
                    "h2d_bytes_per_step": x_host.numel() * 4 + y_host.numel() * 4, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
        }
        line.update(extra)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def _protect_stdout():
    """The contract is ONE JSON line on stdout; libraries (NCCL prints its version banner to stdout, torchrun,
    printf from kernels) must not pollute it: route fd 1 to stderr and keep the real stdout for the result line."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line: dict) -> None:
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="fp16", choices=["fp16", "bf16"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
