#!/usr/bin/env python
"""Benchmark of the ConvLSTM encoder-forecaster hot path (BASELINE.json metric: rollout frames/s, fwd+bwd).

  python bench.py --gpus N --steps K --warmup W              our arm (one process per GPU; torchrun for N>1)
  python bench.py --impl reference --gpus N --steps K ...    the reference's CPU implementation of the same path
  python bench.py --impl eager-gpu                           informational: the same torch ops on the GPU (cuDNN, TF32)
  python bench.py --gpus N --check                           N-GPU sharded gradients == 1-GPU full-batch gradients
  python bench.py --gpus N --global-batch 128                strong-scaling form of BASELINE configs[2] (micro-batches)

One step = zero_grad + forward rollout + MSE loss + backward (BPTT) + gradient all-reduce (N>1) + Adam step
on one batch of synthetic 12-channel sequences, hid 64, 256x256, 12 in / 24 out, batch 16 per GPU (weak scaling;
at N=8 this is BASELINE configs[2] exactly: global batch 128).  frames/s = B * (T_in + T_out) / step time.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import csv
import json
import os
import re
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
# the bounded sample of the workload that the CPU legs time (frames/s is normalised by B * (t_in + t_out))
CPU_SAMPLE = dict(batch=1, t_in=4, t_out=8)


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


def ncu_traffic(kernel_substr: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the named kernel, read from the committed
    `ncu --set full` summary (tools/ncu_summary.py output under profiles/); (None, None) if no capture is committed."""
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for name in ("r2k_ncu_full_summary.csv", "r2h_ncu_full_summary.csv", "r2e_ncu_full_summary.csv", "r2_ncu_full_summary.csv", "r1_v2_ncu_full_summary.csv"):
        path = os.path.join(ROOT, "profiles", name)
        if not os.path.isfile(path):
            continue
        vals = []
        with open(path) as f:
            for row in csv.DictReader(f):
                if kernel_substr not in row.get("kernel", ""):
                    continue
                tot = 0.0
                for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    m = re.match(r"([0-9.]+)\s*(\w+)", row.get(key, ""))
                    if not m:
                        break
                    tot += float(m.group(1)) * unit.get(m.group(2), 1.0)
                else:
                    vals.append(tot)
        if vals:
            return sum(vals) / len(vals), f"profiles/{name} ({len(vals)} launch(es), ncu --set full)"
    return None, None


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
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "power_w_median": statistics.median(pw) if pw else None,
                "power_w_max": max(pw) if pw else None}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_sample(steps: int, warmup: int, threads: int):
    """The reference's CPU fp32 path (torch autograd over cat/Conv2d/sigmoid/tanh, exactly the op sequence of
    layers/ConvLSTM.py:42-57 and conv_lstm.py:171-203, restated in oracle/convlstm_oracle.py because
    /root/reference does not exist on the GPU box).  Bounded sample of the bench workload: same hid/channels/
    256x256, batch 1, 4 in / 8 out steps; frames/s scales linearly in batch and steps."""
    from oracle import convlstm_oracle as O

    torch.set_num_threads(threads)
    B, t_in, t_out = CPU_SAMPLE["batch"], CPU_SAMPLE["t_in"], CPU_SAMPLE["t_out"]
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
    return dict(value=B * (t_in + t_out) / dt, ms_per_step=dt * 1e3, threads=torch.get_num_threads(),
                sample=f"B={B}, {t_in} in / {t_out} out, hid {CFG['hidden']}, {CFG['hw']}x{CFG['hw']}, fwd+bwd+Adam, "
                       f"{steps} timed steps after {warmup} warm-up (torch {torch.__version__} CPU fp32, "
                       f"{torch.get_num_threads()} threads)")


def sample_workload():
    s = CPU_SAMPLE
    return (f"BOUNDED SAMPLE of the workload below, timed per step: batch {s['batch']}, {s['t_in']} in / {s['t_out']} out "
            f"(same hid {CFG['hidden']}, {CFG['channels']}ch, {CFG['hw']}x{CFG['hw']}, fwd+bwd+Adam); frames/s = "
            f"B*(T_in+T_out)/t is per-frame normalised, so it compares with the full configuration: ")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to its workers; this leg is meant to use every host core
    threads = os.cpu_count() or 1
    r = cpu_reference_sample(max(1, args.steps), max(1, min(args.warmup, 1)), threads)
    cfg = workload_config(args.gpus)
    cfg["workload"] = sample_workload() + cfg["workload"]
    cfg["sample"] = dict(CPU_SAMPLE)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": r["value"], "unit": "frames/s", "cores": r["threads"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def cpu_baseline_subprocess():
    """The cpu_baseline leg of our arm: the reference arm run as a child process with a clean threading environment
    (no OMP_NUM_THREADS / MKL_NUM_THREADS inherited from a launcher), one timed step after one warm-up."""
    env = {k: v for k, v in os.environ.items()
           if k not in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "RANK", "LOCAL_RANK", "WORLD_SIZE", "CUDA_VISIBLE_DEVICES")}
    env["CUDA_VISIBLE_DEVICES"] = ""
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                             capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
        d = json.loads([l for l in out.stdout.splitlines() if l.strip()][-1])
        return d["cpu_baseline"]
    except Exception as e:  # the baseline is a reported number, not a dependency of the measurement
        return {"value": None, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}


def workload_config(n, global_batch=None):
    per = CFG["batch_per_gpu"]
    d = {
        "workload": (f"encoder-forecaster ConvLSTM (2 enc + 2 dec cells, 3x3) hid {CFG['hidden']}, {CFG['channels']}ch "
                     f"{CFG['hw']}x{CFG['hw']}, {CFG['t_in']} in / {CFG['t_out']} out, batch {per}/GPU, "
                     "fwd+bwd+grad-allreduce+Adam (BASELINE configs[2]; shape of configs[1])"),
        "global_batch": per * n, "parallelism": f"dp{n}",
        "l2": "working set (67 GB workspace of saved states, ~335 GB of HBM traffic per step) >> 126 MB L2; no flush needed",
    }
    if global_batch:
        d["global_batch"] = global_batch
        d["micro_batches_per_gpu"] = global_batch // (n * per)
        d["workload"] = d["workload"].replace(f"batch {per}/GPU", f"global batch {global_batch} = {global_batch // n}/GPU in "
                                              f"micro-batches of {per} with gradient accumulation")
    return d


# ------------------------------------------------------------------------------------------ informational GPU-eager arm
def run_eager_gpu(args):
    """SURVEY.md §0 names PyTorch eager (cuDNN conv + ATen pointwise) on the same B200 as the bar to beat: the
    reference's op sequence (layers/ConvLSTM.py:42-57 inside the loops of conv_lstm.py:171-203) written with plain
    torch modules on the GPU, torch's defaults for convolutions (cudnn.allow_tf32 = True), autograd backward, fused
    Adam, largest batch of {16, 8, 4, 2, 1} that fits next to autograd's saved activations.  Informational; not part of
    the driver's contract and not the parity oracle."""
    import torch.nn as nn

    class Cell(nn.Module):
        def __init__(self, cin, hid):
            super().__init__()
            self.hid = hid
            self.conv = nn.Conv2d(cin + hid, 4 * hid, 3, padding=1, bias=True)

        def forward(self, x, h, c):
            i, f, o, g = torch.split(self.conv(torch.cat([x, h], dim=1)), self.hid, dim=1)
            i, f, o, g = torch.sigmoid(i), torch.sigmoid(f), torch.sigmoid(o), torch.tanh(g)
            c = f * c + i * g
            return o * torch.tanh(c), c

    class Net(nn.Module):
        def __init__(self, cin, hid, cout):
            super().__init__()
            self.hid = hid
            self.cells = nn.ModuleList([Cell(cin, hid), Cell(hid, hid), Cell(hid, hid), Cell(hid, hid)])
            self.head = nn.Conv3d(hid, cout, (1, 3, 3), padding=(0, 1, 1))

        def forward(self, x, t_out):
            B, T, _, H, W = x.shape
            hs = [torch.zeros(B, self.hid, H, W, device=x.device) for _ in range(4)]
            cs = [torch.zeros(B, self.hid, H, W, device=x.device) for _ in range(4)]
            for t in range(T):
                hs[0], cs[0] = self.cells[0](x[:, t], hs[0], cs[0])
                hs[1], cs[1] = self.cells[1](hs[0], hs[1], cs[1])
            vec, outs = hs[1], []
            for _ in range(t_out):
                hs[2], cs[2] = self.cells[2](vec, hs[2], cs[2])
                hs[3], cs[3] = self.cells[3](hs[2], hs[3], cs[3])
                vec = hs[3]
                outs.append(vec)
            return torch.sigmoid(self.head(torch.stack(outs, 1).permute(0, 2, 1, 3, 4)))

    dev = torch.device("cuda", 0)
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    t_in, t_out = CFG["t_in"], CFG["t_out"]
    res, tried = None, []
    for B in (16, 8, 4, 2, 1):
        net = x = tgt = opt = None
        try:
            torch.manual_seed(0)
            net = Net(CFG["channels"], CFG["hidden"], CFG["out_channels"]).to(dev)
            x = torch.randn(B, t_in, CFG["channels"], CFG["hw"], CFG["hw"], device=dev)
            tgt = torch.rand(B, t_out, CFG["out_channels"], CFG["hw"], CFG["hw"], device=dev)
            opt = torch.optim.Adam(net.parameters(), lr=1e-4, fused=True)

            def step():
                opt.zero_grad(set_to_none=True)
                loss = torch.nn.functional.mse_loss(net(x, t_out).permute(0, 2, 1, 3, 4), tgt)
                loss.backward()
                opt.step()

            for _ in range(max(2, args.warmup)):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            with torch.no_grad():
                for _ in range(2):
                    net(x, t_out)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(args.steps):
                    net(x, t_out)
                e1.record()
                torch.cuda.synchronize()
            ms_inf = e0.elapsed_time(e1) / args.steps
            res = dict(batch=B, ms_per_step=ms, frames_per_s=B * (t_in + t_out) / ms * 1e3,
                       inference_ms=ms_inf, inference_frames_per_s=B * (t_in + t_out) / ms_inf * 1e3,
                       peak_mem_gb=torch.cuda.max_memory_allocated() / 1e9, batches_that_did_not_fit=tried)
            break
        except torch.OutOfMemoryError:
            tried.append(B)
            del net, x, tgt, opt
            torch.cuda.empty_cache()
    emit({"impl": "eager-gpu", "metric": METRIC, "value": res["frames_per_s"] if res else None, "unit": "frames/s",
          "n_gpus": 1, "dtype": "tf32 conv (cudnn.allow_tf32) / fp32 pointwise", "data": "synthetic",
          "config": workload_config(1), "result": res,
          "note": "torch eager: cat -> cuDNN Conv2d -> split -> sigmoid/tanh -> ... with autograd, fused Adam"})


# ------------------------------------------------------------------------------------------ our arm
def setup_dist(args):
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit(f"--gpus {args.gpus} needs torchrun (--nproc-per-node {args.gpus})")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    return dist, world, rank, local, dev


def run_check(args):
    """Multi-GPU correctness on hardware: (1) gradients of a global batch sharded over N ranks and
    averaged by the one all-reduce == the gradients one GPU computes on the whole batch (rel-L2 <= 2e-3 per tensor);
    (2) after the Adam step every replica holds bit-identical parameters."""
    from satflow_b200 import EncoderDecoderConvLSTM
    from satflow_b200.distributed import FlatGradBucket, broadcast_parameters, shard_batch

    dist, world, rank, local, dev = setup_dist(args)
    per, t_in, t_out, hw = 2, 4, 6, 64
    G = per * world
    torch.manual_seed(0)
    model = EncoderDecoderConvLSTM(hidden_dim=64, input_channels=12, out_channels=12, forecast_steps=t_out, lr=1e-3).to(dev)
    broadcast_parameters(model)
    g = torch.Generator().manual_seed(99)
    x_all = torch.randn(G, t_in, 12, hw, hw, generator=g).to(dev)
    y_all = torch.rand(G, t_out, 12, hw, hw, generator=g).to(dev)
    bucket = FlatGradBucket(model.parameters())
    # 1-GPU full-batch gradients (every rank computes them: same data, same weights)
    bucket.zero_()
    model.training_step((x_all, y_all), 0).backward()
    full = bucket.flat.clone()
    # sharded
    bucket.zero_()
    model.training_step((shard_batch(x_all, rank, world).contiguous(), shard_batch(y_all, rank, world).contiguous()), 0).backward()
    bucket.all_reduce_mean()
    worst, off = 0.0, 0
    for p_ in bucket.params:
        n = p_.numel()
        a, b = bucket.flat[off:off + n], full[off:off + n]
        worst = max(worst, float((a - b).norm() / b.norm().clamp_min(1e-30)))
        off += n
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused=True)
    opt.step()
    flat_p = torch.cat([p_.detach().reshape(-1) for p_ in model.parameters()])
    ident = True
    if world > 1:
        ref = flat_p.clone()
        dist.broadcast(ref, src=0)
        same = torch.tensor([1.0 if torch.equal(ref, flat_p) else 0.0], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        ident = bool(same.item() == 1.0)
        w = torch.tensor([worst], device=dev)
        dist.all_reduce(w, op=dist.ReduceOp.MAX)
        worst = float(w.item())
    ok = worst <= 2e-3 and ident
    if rank == 0:
        emit({"check": "multi-gpu gradient equivalence", "n_gpus": world, "ok": ok, "worst_rel_l2_vs_full_batch": worst,
              "replicas_bit_identical_after_adam": ident, "tolerance": 2e-3,
              "config": {"workload": f"hid 64, 12ch {hw}x{hw}, {t_in} in / {t_out} out, global batch {G} = {per}/GPU"}})
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if not ok:
        raise SystemExit(1)


def parse_trace(report: str):
    """clstm_trace_report table -> {kernel name: (count, total ms, avg us)}."""
    out = {}
    for ln in report.splitlines():
        m = re.match(r"\s*([0-9.]+) ms\s+[0-9.]+%\s+n=\s*(\d+)\s+avg=\s*([0-9.]+) us\s+(.*)$", ln)
        if m:
            out[m.group(4).strip()] = (int(m.group(2)), float(m.group(1)), float(m.group(3)))
    return out


def run_ours(args):
    from satflow_b200 import ConvLSTM, ConvLSTMCell, EncoderDecoderConvLSTM, _lib
    from satflow_b200.distributed import FlatGradBucket, broadcast_parameters

    dist, world, rank, local, dev = setup_dist(args)
    L = _lib.lib()
    _lib.check(L.clstm_device_check(local))

    B, t_in, t_out = CFG["batch_per_gpu"], CFG["t_in"], CFG["t_out"]
    C, Co, hid, HW = CFG["channels"], CFG["out_channels"], CFG["hidden"], CFG["hw"]
    n_micro = 1
    if args.global_batch:
        if args.global_batch % (world * B):
            raise SystemExit(f"--global-batch must be a multiple of {world * B}")
        n_micro = args.global_batch // (world * B)
    torch.manual_seed(0)
    model = EncoderDecoderConvLSTM(hidden_dim=hid, input_channels=C, out_channels=Co, forecast_steps=t_out, lr=1e-4)
    model.model.operand_dtype = args.dtype
    model = model.to(dev)
    broadcast_parameters(model)
    bucket = FlatGradBucket(model.parameters())
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)
    g = torch.Generator().manual_seed(1234 + rank)
    x_host = torch.randn(n_micro * B, t_in, C, HW, HW, generator=g).pin_memory()
    y_host = torch.rand(n_micro * B, t_out, Co, HW, HW, generator=g).pin_memory()
    x_dev, y_dev = x_host.to(dev), y_host.to(dev)

    def train_step(x, tgt):
        bucket.zero_()
        loss = None
        for m in range(n_micro):  # micro-batches of 16 accumulate into .grad (autograd's +=); mean over the local batch
            xm, tm = x[m * B:(m + 1) * B], tgt[m * B:(m + 1) * B]
            loss = model.training_step((xm, tm), 0)  # conv_lstm.py:53-70: forward, MSE loss, per-frame losses
            (loss / n_micro if n_micro > 1 else loss).backward()
        bucket.all_reduce_mean()
        opt.step()
        return loss

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

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
        return max_over_ranks(e0.elapsed_time(e1)) / steps, (L.clstm_launch_count() - n0)

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

    loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()
    loss_ev = [torch.cuda.Event(), torch.cuda.Event()]
    n_e2e_warm = 2
    # ONE loader pipeline across warm-up and timed steps (steady state): while step i runs, the copy of step i + 1's
    # inputs is in flight.  The timed region therefore holds exactly one H2D copy per timed step (those of steps
    # 1 .. K and of one spare batch that keeps the pipeline full), and every step's loss is read on the host — one step
    # late, from a pinned two-slot ring guarded by events — so the host stays one step ahead of the device, as a
    # training loop that logs its loss does.
    pf = iter(DevicePrefetcher(((x_host, y_host) for _ in range(n_e2e_warm + args.steps + 1)), dev))
    losses = []

    def e2e_steps(n, i0):
        for i in range(i0, i0 + n):
            xd, td = next(pf)
            loss = train_step(xd, td)
            pf.done_with_current()
            loss_host[i & 1].copy_(loss.detach(), non_blocking=True)
            loss_ev[i & 1].record()
            if i > i0:
                loss_ev[(i - 1) & 1].synchronize()
                losses.append(float(loss_host[(i - 1) & 1]))
        loss_ev[(i0 + n - 1) & 1].synchronize()
        losses.append(float(loss_host[(i0 + n - 1) & 1]))

    e2e_steps(n_e2e_warm, 0)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_steps(args.steps, n_e2e_warm)
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    assert len(losses) == n_e2e_warm + args.steps and all(v == v for v in losses)

    # inference (BASELINE configs[1]): forward rollout only, same shapes
    extra = {}

    def infer():
        with torch.no_grad():
            model(x_dev[:B], t_out)

    for _ in range(3):
        infer()
    ms_inf, _ = timed(infer, args.steps)
    extra["inference"] = {"frames_per_s": world * B * (t_in + t_out) / ms_inf * 1e3, "ms_per_step": ms_inf,
                          "tflops": algorithmic_flops_fwd(B, t_in, t_out, C, hid, Co, HW, HW) / ms_inf / 1e9,
                          "config": "BASELINE configs[1]: fwd only, batch 16 per GPU"}

    if world == 1 and not args.no_extras and args.dtype == "fp16":
        # informational: the same inference rollout with bf16 operands (opt-in; outputs within 2e-3, recurrent states not —
        # DESIGN.md §2).  Faster under the power cap: fewer multiplier bits toggle.
        prev = model.model.operand_dtype
        try:
            model.model.operand_dtype = "bf16"
            for _ in range(3):
                infer()
            ms_b, _ = timed(infer, args.steps)
            extra["inference"]["bf16_operands"] = {"frames_per_s": B * (t_in + t_out) / ms_b * 1e3, "ms_per_step": ms_b}
        except Exception as e:  # informational leg: never costs the result line
            extra["inference"]["bf16_operands"] = {"error": str(e)[:200]}
        finally:
            model.model.operand_dtype = prev

    peaks = measured_peaks()
    roof = None
    cpu = None
    # (1) IN SITU: one CUDA event after every library launch of two more training steps (clstm_trace_enable), i.e.
    # each kernel's average duration under the clocks and cache state of the real step -> sustained peak.  Every rank
    # runs the steps (they contain the gradient all-reduce); only rank 0 records and reports.
    if rank == 0:
        _lib.trace_enable(4096 * n_micro)
    for _ in range(2):
        train_step(x_dev, y_dev)
    barrier()
    if rank == 0:
        tr = parse_trace(_lib.trace_report())
        _lib.trace_enable(0)
        # the fused dgrad launches that also carry the head's dgrad as a second K segment are the same kernel
        hk, fk = "dgradT_fused2_kernel[+head dgrad]", "dgradT_fused2_kernel"
        if hk in tr:
            extra["fused_dgrad_with_head_segment"] = {"n_per_step": tr[hk][0] // 2, "avg_us": tr[hk][2],
                                                      "plain_avg_us": tr.get(fk, (0, 0.0, None))[2]}
            n0, ms0, _ = tr.get(fk, (0, 0.0, 0.0))
            tr[fk] = (n0 + tr[hk][0], ms0 + tr[hk][1], (ms0 + tr[hk][1]) * 1e3 / (n0 + tr[hk][0]))
            del tr[hk]
        fl = cell_step_flops(B, hid, hid, HW, HW)  # a cell step with a 64-channel input: K = (64 + 64) * 9
        npix = B * HW * HW
        gg_bytes = npix * hid * (4 * 2 + 4 + 4 + 2 * 4 + 2 * 4 + 4 * 2)  # gates, c_prev, c_next, 2 dh, dc r/w, dz
        # algorithmic bytes per launch (16-bit x / h / gates / dz, fp32 c / dc / dh; weights negligible)
        px = npix * hid
        c16 = args.dtype == "fp16" and os.environ.get("CLSTM_C16", "1") != "0"   # c crosses HBM in 16 bits
        cb = 2 if c16 else 4
        by_cell = px * (2 + 2 + cb + 2 + cb + 8)          # x, h_prev, c_prev read; h, c, 4 gates written
        rc = args.dtype == "fp16" and os.environ.get("CLSTM_RECOMP_C", "1") != "0"
        s16 = rc and os.environ.get("CLSTM_STATE16", "1") != "0"   # dh_prev / own dh / dc in 16 bits
        by_fused = px * (8 + (2 if s16 else 4) + 8 + cb + (0 if rc else cb) + (2 if s16 else 4) + (4 if s16 else 8) + 8)  # dz read, dh_prev written; gates, c_prev, (c_next
                                                           # unless it is recomputed from the gates), ONE dh source from HBM
                                                           # (the other is this launch's dx, in shared memory; with the head
                                                           # segment the head's G tile replaces dstack, same bytes), dc r/w,
                                                           # dz written
        by_wgrad = px * (8 + 2 + 2)                        # dz, x, h_prev read
        by_dgrad = px * (8 + 4 + 4)                        # dz read, dx and dh_prev written
        tensor_kernels = {
            "cell_step": ("convgemm_kernel<EPI_LSTM>: fused conv + LSTM cell step (training variant, also writes gates)",
                          "convgemm_kernel<__half, 0>", by_cell),
            "cell_step[pair]": ("cellstep_pair_kernel: fused conv + LSTM cell step on CTA pairs (cta_group::2; training "
                                "variant, also writes gates)", "cellstep_pair_kernel", by_cell),
            "dgradT_fused2_kernel": ("dgradT_fused2_kernel: data gradient + gate gradient of the next chain step on "
                                     "dedicated worker warps", "dgradT_fused", by_fused),
            "dgradT_fused_kernel": ("dgradT_fused_kernel: data gradient + fused gate gradient of the next chain step",
                                    "dgradT_fused", by_fused),
            "wgrad[halo rows]": ("wgrad_kernel (halo rows): weight gradient", "wgrad_kernel", by_wgrad),
            "wgrad_gate_kernel": ("wgrad_kernel + gate-gradient worker warps", "wgrad_kernel", by_wgrad + gg_bytes),
            "dgradT_kernel": ("dgradT_kernel: data gradient", "dgradT_kernel", by_dgrad),
        }
        in_situ = []
        for key, (label, ncu_name, nbytes) in tensor_kernels.items():
            if key not in tr:
                continue
            n, tot_ms, avg_us = tr[key]
            sec = avg_us * 1e-6
            t_tensor = fl / (peaks["bf16_sustained"] * 1e12)   # roofline times of this launch on either resource
            t_hbm = nbytes / (peaks["hbm"] * 1e9)
            traffic, tsrc = ncu_traffic(ncu_name)
            common = {"kernel": label, "traffic": traffic, "traffic_source": tsrc, "launch_ms": avg_us * 1e-3,
                      "launches_per_step": n // 2, "share_of_step_ms": tot_ms / 2, "flops_per_launch": fl,
                      "bytes_per_launch": nbytes,
                      "tensor": {"achieved": fl / sec / 1e12, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                                 "frac": t_tensor / sec},
                      "hbm": {"achieved": nbytes / sec / 1e9, "peak": peaks["hbm"], "unit": "GB/s", "frac": t_hbm / sec},
                      "peak_source": peaks["source"] + "; tensor: sustained figure (kernel timed inside the training "
                                     "step), hbm: the copy bandwidth",
                      "timing": "CUDA event after every launch of 2 training steps on the launching stream "
                                "(clstm_trace_enable), taken right after the timed region"}
            # the binding resource is the one whose roofline time for this launch is longer (arithmetic intensity
            # against the ridge point): the fused dgrad carries 3.2 GB for 0.62 TFLOP = 191 FLOP/B < 212
            if t_hbm > t_tensor:
                common.update({"bound": "hbm", "achieved": nbytes / sec / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
                               "frac": t_hbm / sec})
            else:
                common.update({"bound": "tensor", "achieved": fl / sec / 1e12, "peak": peaks["bf16_sustained"],
                               "unit": "TFLOP/s", "frac": t_tensor / sec})
            in_situ.append(common)
        if "gate_grad_kernel" in tr:
            n, tot_ms, avg_us = tr["gate_grad_kernel"]
            in_situ.append({"kernel": "gate_grad_kernel (pointwise gate gradient, standalone launches)", "bound": "hbm",
                            "achieved": gg_bytes / (avg_us * 1e-6) / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
                            "frac": gg_bytes / (avg_us * 1e-6) / 1e9 / peaks["hbm"], "launch_ms": avg_us * 1e-3,
                            "launches_per_step": n // 2, "share_of_step_ms": tot_ms / 2})
        in_situ.sort(key=lambda d: -d.get("share_of_step_ms", 0.0))
        if in_situ:
            roof = in_situ[0]  # the kernel with the largest share of the step
            extra["kernels_in_step"] = in_situ[1:]
        extra["trace_top"] = {k: {"n_per_step": v[0] // 2, "ms_per_step": v[1] / 2, "avg_us": v[2]}
                              for k, v in sorted(tr.items(), key=lambda kv: -kv[1][1])[:10]}

        # (2) ISOLATED: each kernel launched alone, back to back, through the C-ABI measurement hook -> burst peak.
        # (20 back-to-back launches of one tensor-bound kernel draw more power than the mixed step: lower clocks.)
        plan = [p for p in model.model._plans.values() if p.training and p.cfg.dtype == _lib.DTYPES[args.dtype]][0]

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

        iso = []
        for kind, name in (("cell_fwd", "fused cell step"), ("dgrad_fused", "dgradT + fused gate gradient"),
                           ("dgrad", "dgradT alone"), ("wgrad", "wgrad")):
            try:
                ms = time_kernel(kind, 3, 5)
            except Exception as e:
                iso.append({"kernel": name, "error": str(e)[:120]})
                continue
            iso.append({"kernel": name, "bound": "tensor", "achieved": fl / ms / 1e9, "peak": peaks["bf16_burst"],
                        "unit": "TFLOP/s", "frac": fl / ms / 1e9 / peaks["bf16_burst"], "launch_ms": ms})
        extra["kernels_isolated"] = {"how": "20 back-to-back launches of one kernel on the plan's tensors (timing only: "
                                            "backward kinds re-run on whatever the last backward left), burst peak",
                                     "rows": iso}
        flops_step = 3 * algorithmic_flops_fwd(B * n_micro, t_in, t_out, C, hid, Co, HW, HW)
        extra["step_tflops"] = flops_step * world / ms_step / 1e9
        extra["step_frac_of_sustained_peak"] = flops_step / ms_step / 1e9 / peaks["bf16_sustained"]
        # Energy view (DESIGN.md §4 "board power"): every GEMM kernel of the step, like cuBLAS, runs at the board's power
        # cap, so joules per step is what the step time follows.  Power = median of the nvidia-smi samples taken during
        # the timed region on rank 0 (few samples: indicative; tools/energy_probe.py is the per-kernel measurement).
        try:
            pw = (clocks or {}).get("power_w_median")
            if pw:
                extra["energy"] = {"board_power_w_median": pw, "joule_per_step": pw * ms_step * 1e-3,
                                   "pj_per_algorithmic_flop": pw * ms_step * 1e-3 / flops_step * 1e12,
                                   "cublas_bf16_pj_per_flop_same_pool": 0.709,
                                   "source": "clocks sampler; cuBLAS figure from profiles/r2c_energy_probe.json"}
        except Exception:
            pass
        try:
            st = model.model.check_gradients()
            if st:
                extra["grad_range"] = {k: st[k] for k in ("scale", "amax_dlogit", "amax_dz_scaled", "headroom_log2")}
        except FloatingPointError as e:
            extra["grad_range"] = {"error": str(e)[:200]}

    # other BASELINE configs (rank 0, single GPU, on request off): parity for them lives in tests/; here only numbers
    if rank == 0 and world == 1 and not args.no_extras:
        model.model.release_plans()
        del model, bucket, opt, x_dev, y_dev
        torch.cuda.empty_cache()
        extra["other_configs"] = other_configs(ConvLSTM, ConvLSTMCell, dev, peaks)

    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline_subprocess()

    if rank == 0:
        frames = world * B * n_micro * (t_in + t_out)
        line = {
            "metric": METRIC, "value": frames / ms_step * 1e3, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong" if args.global_batch else "weak",
            "vs_baseline": None, "dtype": "f16" if args.dtype == "fp16" else "bf16", "data": "synthetic",
            "config": workload_config(world, args.global_batch),
            "e2e": {"value": frames / ms_e2e * 1e3, "unit": "frames/s",
                    "h2d_bytes_per_step": x_host.numel() * 4 + y_host.numel() * 4,
                    "d2h_bytes_per_step": 4 + 16,  # the loss + the gradient range statistics of the step
                    "ms_per_step": ms_e2e,
                    "how": "module API (training_step, backward, all-reduce, Adam) fed by satflow_b200.prefetch."
                           "DevicePrefetcher from PINNED HOST tensors: one cudaMemcpyAsync H2D of x and target per step "
                           "on a side stream, double-buffered, so the copy of step i+1 overlaps step i (33 ms of PCIe "
                           "under 165 ms of compute); the pipeline runs continuously through 2 warm-up and the K timed "
                           "steps, so K copies fall inside the timed region; every step's loss is copied D2H and read "
                           "on the host one step later"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
        }
        line.update(extra)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def other_configs(ConvLSTM, ConvLSTMCell, dev, peaks):
    """BASELINE configs[0], [3] and a subset of the configs[4] cell sweep, device-timed (CUDA events, 3 warm-up)."""
    out = {}

    def timeit(fn, reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    def rollout(name, B, t_in, t_out, hid, hw, n_layers, reps):
        torch.manual_seed(0)
        net = ConvLSTM(12, hid, 12, n_layers=n_layers).to(dev)
        x = torch.randn(B, t_in, 12, hw, hw, device=dev)
        tgt = torch.rand(B, t_out, 12, hw, hw, device=dev)

        def train():
            net.zero_grad(set_to_none=True)
            torch.nn.functional.mse_loss(net(x, t_out).permute(0, 2, 1, 3, 4), tgt).backward()

        def infer():
            with torch.no_grad():
                net(x, t_out)

        try:
            ms_t, ms_i = timeit(train, reps), timeit(infer, reps)
            fl = algorithmic_flops_fwd(B, t_in, t_out, 12, hid, 12, hw, hw, n_layers=n_layers)
            out[name] = {"fwd_bwd_ms": ms_t, "fwd_bwd_frames_per_s": B * (t_in + t_out) / ms_t * 1e3,
                         "fwd_bwd_tflops": 3 * fl / ms_t / 1e9, "fwd_ms": ms_i,
                         "fwd_frames_per_s": B * (t_in + t_out) / ms_i * 1e3, "fwd_tflops": fl / ms_i / 1e9}
        except Exception as e:
            out[name] = {"error": str(e)[:200]}
        net.release_plans()
        del net, x, tgt
        torch.cuda.empty_cache()

    rollout("configs[0]: 1-layer hid 32, 12ch 64x64, 4 in / 4 out, batch 2", 2, 4, 4, 32, 64, 1, 20)
    rollout("configs[0] on the reference's 2+2 architecture", 2, 4, 4, 32, 64, 2, 20)
    rollout("configs[3]: 3-layer hid 128, 12ch 512x512, 12 in / 12 out, batch 1", 1, 12, 12, 128, 512, 3, 3)
    cells = []
    for k, hid, px in ((3, 64, 128), (3, 64, 256), (3, 64, 512), (5, 64, 256), (3, 128, 256), (3, 256, 256), (5, 128, 128)):
        try:
            torch.manual_seed(0)
            cell = ConvLSTMCell(hid, hid, (k, k), True).to(dev)
            x = torch.randn(1, hid, px, px, device=dev)
            h = torch.randn(1, hid, px, px, device=dev) * 0.5
            c = torch.randn(1, hid, px, px, device=dev)

            def fwd():
                with torch.no_grad():
                    cell(x, [h, c])

            hg, cg = h.clone().requires_grad_(True), c.clone().requires_grad_(True)

            def fwd_bwd():
                cell.zero_grad(set_to_none=True)
                hn, cn = cell(x, [hg, cg])
                (hn.sum() + cn.sum()).backward()

            fl = cell_step_flops(1, hid, hid, px, px, k)
            ms_f, ms_fb = timeit(fwd, 10), timeit(fwd_bwd, 5)
            stepper = cell.native(1, (px, px)).reset(h, c).set_input(x)  # state kept in the device layout
            ms_n = timeit(stepper.step, 20)
            stepper.close()
            cells.append({"kernel": f"{k}x{k}", "hidden": hid, "px": px, "fwd_ms": ms_f, "fwd_tflops": fl / ms_f / 1e9,
                          "fwd_bwd_ms": ms_fb, "fwd_bwd_tflops": 3 * fl / ms_fb / 1e9,
                          "native_step_ms": ms_n, "native_step_tflops": fl / ms_n / 1e9})
            del cell, x, h, c, hg, cg
            torch.cuda.empty_cache()
        except Exception as e:
            cells.append({"kernel": f"{k}x{k}", "hidden": hid, "px": px, "error": str(e)[:200]})
    out["configs[4] subset: ConvLSTMCell, B=1, Cin_x = hidden; fwd / fwd_bwd through the drop-in NCHW fp32 module API "
        "(layout conversion on every call), native_step = ConvLSTMCell.native() (state kept in the device layout)"] = cells
    return out


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
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "eager-gpu"])
    ap.add_argument("--dtype", default="fp16", choices=["fp16", "bf16"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the other BASELINE configs (configs[0], [3], [4])")
    ap.add_argument("--check", action="store_true", help="multi-GPU gradient-equivalence check instead of the benchmark")
    ap.add_argument("--global-batch", type=int, default=0,
                    help="strong scaling: fixed global batch (multiple of 16 * gpus), micro-batches of 16 per GPU")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "eager-gpu":
        run_eager_gpu(args)
    elif args.check:
        run_check(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
