"""BASELINE.json configs[3] and configs[4] on one B200, with the reference's CPU path (oracle port) beside it.

configs[4]: ConvLSTMCell microbench sweep — kernel 3x3/5x5, hidden 64/128/256, 128-512 px (B=1), forward and
forward+backward through the drop-in ConvLSTMCell (C ABI clstm_cell_forward/backward; includes the NCHW<->NHWC
pack/unpack kernels), vs the CPU fp32 oracle on all host cores where one call takes < ~10 s.
configs[3]: 3-layer ConvLSTM hidden 128, 12ch 512x512, 12 in / 12 out, B=1: rollout forward and training step.
Writes gpurun_out/sweep.json and a markdown table on stdout."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from satflow_b200 import ConvLSTM, ConvLSTMCell
from oracle import convlstm_oracle as O


def gpu_time(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def cpu_time(fn, iters=2, warm=1):
    for _ in range(warm):
        fn()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    return (time.perf_counter() - t0) / iters * 1e3


def main():
    torch.set_num_threads(os.cpu_count())
    rows = []
    quick = "--quick" in sys.argv
    for k in (3, 5):
        for hid in (64, 128, 256):
            for px in (128, 256, 512):
                for cin in (12, hid):
                    if quick and (px == 512 or hid == 256):
                        continue
                    fl = 2 * px * px * (cin + hid) * 4 * hid * k * k
                    cell = ConvLSTMCell(cin, hid, (k, k), True).cuda()
                    x = torch.randn(1, cin, px, px, device="cuda")
                    h = torch.randn(1, hid, px, px, device="cuda") * 0.5
                    c = torch.randn(1, hid, px, px, device="cuda")

                    def fwd():
                        with torch.no_grad():
                            cell(x, [h, c])

                    def fwdbwd():
                        xx = x.clone().requires_grad_(True)
                        hn, cn = cell(xx, [h, c])
                        (hn.sum() + cn.sum()).backward()

                    t_f, t_fb = gpu_time(fwd), gpu_time(fwdbwd)
                    row = {"k": k, "hid": hid, "px": px, "cin": cin, "gflop_fwd": fl / 1e9, "gpu_fwd_ms": t_f,
                           "gpu_fwdbwd_ms": t_fb, "gpu_fwd_tflops": fl / t_f / 1e9, "gpu_fwdbwd_tflops": 3 * fl / t_fb / 1e9}
                    if fl < 2.0e11:  # CPU leg only where one call stays well below ~10 s
                        w = cell.conv.weight.detach().cpu()
                        b_ = cell.conv.bias.detach().cpu()
                        xc, hc, cc = x.cpu(), h.cpu(), c.cpu()

                        def cfwd():
                            with torch.no_grad():
                                O.cell_forward(xc, hc, cc, w, b_)

                        def cfwdbwd():
                            wr = w.clone().requires_grad_(True)
                            hn, cn, _ = O.cell_forward(xc, hc, cc, wr, b_)
                            (hn.sum() + cn.sum()).backward()

                        row["cpu_fwd_ms"] = cpu_time(cfwd)
                        row["cpu_fwdbwd_ms"] = cpu_time(cfwdbwd)
                    rows.append(row)
                    for p in list(cell._plans.values()):
                        p.close()
                    cell._plans.clear()
                    print(json.dumps(row), flush=True)
    # configs[3]
    cfg3 = {}
    if not quick:
        torch.manual_seed(0)
        net = ConvLSTM(12, 128, 12, n_layers=3).cuda()
        x = torch.randn(1, 12, 12, 512, 512, device="cuda")
        tgt = torch.rand(1, 12, 12, 512, 512, device="cuda")

        def infer():
            with torch.no_grad():
                net(x, 12)

        def train():
            net.zero_grad(set_to_none=True)
            y = net(x, 12)
            torch.nn.functional.mse_loss(y.permute(0, 2, 1, 3, 4), tgt).backward()

        fl = 0
        for cidx in range(6):
            cx = 12 if cidx == 0 else 128
            fl += 2 * 512 * 512 * (cx + 128) * 4 * 128 * 9 * 12
        fl += 2 * 12 * 512 * 512 * 128 * 12 * 9
        t_i = gpu_time(infer, 3, 2)
        t_t = gpu_time(train, 3, 2)
        cfg3 = {"config": "3-layer ConvLSTM hid 128, 12ch 512x512, 12 in / 12 out, B=1", "tflop_fwd": fl / 1e12,
                "infer_ms": t_i, "infer_frames_per_s": 24 / t_i * 1e3, "infer_tflops": fl / t_i / 1e9,
                "train_ms": t_t, "train_frames_per_s": 24 / t_t * 1e3, "train_tflops": 3 * fl / t_t / 1e9,
                "training_workspace_gib": [p.workspace_bytes / 2**30 for p in net._plans.values() if p.training]}
        print(json.dumps(cfg3), flush=True)
        net.release_plans()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"cores": os.cpu_count(), "cells": rows, "cfg3": cfg3}, open(os.path.join(ROOT, "gpurun_out", "sweep.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
