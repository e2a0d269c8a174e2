"""Diagnostic runner for a GPU box: runs every parity stage, prints per-tensor rel-L2, never stops at the
first failure.  Usage:  python tools/gpu_check.py [--quick]   (writes gpurun_out/check.json)"""
import json
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402


def main():
    import gpu_checks as G
    from satflow_b200 import _lib

    results = {}
    stages = []

    def stage(name, fn):
        t0 = time.time()
        try:
            r = fn()
            bad = {k: v for k, v in r.items() if not (v <= G.TOL)}
            results[name] = r
            print(f"[{'OK ' if not bad else 'BAD'}] {name} ({time.time() - t0:.1f}s): " +
                  ", ".join(f"{k}={v:.2e}" for k, v in r.items()), flush=True)
        except Exception:
            results[name] = {"error": traceback.format_exc()}
            print(f"[ERR] {name}:\n{traceback.format_exc()}", flush=True)
        torch.cuda.synchronize()

    def selftest():
        out = torch.zeros(2 * 9, device="cuda")
        _lib.check(_lib.lib().clstm_selftest_shifted_desc(_lib.ptr(out), 2, 9, None))
        torch.cuda.synchronize()
        m = out.cpu().view(2, 9)
        print("shifted-descriptor max|err| (rows: base_offset=0 / (addr>>7)&7; cols: row shift 0..8):")
        print(m)
        return {f"v{v}s{s}": float(m[v, s]) for v in range(2) for s in range(9)}

    stage("selftest_shifted_desc", selftest)
    stage("cell_fwd_small_fp16", lambda: G.cell_case(2, 12, 32, 16, 16, 3, 3, "fp16", backward=False))
    stage("cell_fwd_small_bf16", lambda: {k: v / 8 for k, v in G.cell_case(2, 12, 32, 16, 16, 3, 3, "bf16", backward=False).items()})
    stage("cell_fwdbwd_small_fp16", lambda: G.cell_case(2, 12, 32, 16, 16, 3, 3, "fp16"))
    stage("cell_fwdbwd_odd_fp16", lambda: G.cell_case(1, 5, 8, 7, 9, 3, 5, "fp16"))
    stage("cell_fwdbwd_h64_fp16", lambda: G.cell_case(1, 64, 64, 64, 64, 3, 3, "fp16"))
    stage("cell_fwdbwd_h128_k5_fp16", lambda: G.cell_case(1, 12, 128, 32, 32, 5, 5, "fp16"))
    stage("cell_pair_W256", lambda: G.cell_case(1, 12, 64, 20, 256, 3, 3, "fp16"))
    stage("cell_pair_W200_B2", lambda: G.cell_case(2, 64, 64, 6, 200, 3, 3, "fp16"))
    stage("cell_pair_odd_tiles", lambda: G.cell_case(1, 12, 32, 3, 128, 3, 3, "fp16"))
    stage("cell_pair_k5", lambda: G.cell_case(1, 12, 64, 9, 256, 5, 5, "fp16"))
    stage("cell_pair_h128", lambda: G.cell_case(1, 12, 128, 5, 384, 3, 3, "fp16"))
    stage("rollout_pair_W256", lambda: G.rollout_case(1, 3, 3, 12, 64, 12, 6, 256))
    stage("cell_golden_k3", lambda: G.cell_golden("cell_k3"))
    stage("cell_golden_k35", lambda: G.cell_golden("cell_k35"))
    stage("rollout_fwd_cfg1arch", lambda: G.rollout_case(2, 4, 4, 12, 32, 12, 64, 64, backward=False))
    stage("rollout_fwdbwd_small", lambda: G.rollout_case(2, 3, 4, 12, 16, 5, 12, 10))
    stage("rollout_fwdbwd_cfg1arch", lambda: G.rollout_case(2, 4, 4, 12, 32, 12, 64, 64))
    stage("rollout_fwdbwd_L1", lambda: G.rollout_case(2, 4, 4, 12, 32, 12, 64, 64, n_layers=1))
    stage("rollout_fwdbwd_stress_x3", lambda: G.rollout_case(2, 6, 8, 12, 64, 12, 32, 32, weight_scale=3.0))
    stage("rollout_golden_h16", lambda: G.rollout_golden("rollout_h16_12x10"))
    stage("rollout_golden_h8_stress", lambda: G.rollout_golden("rollout_h8_stress"))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "check.json"), "w") as f:
        json.dump(results, f, indent=1)
    print("launches:", _lib.lib().clstm_launch_count())


if __name__ == "__main__":
    main()
