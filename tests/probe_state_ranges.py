"""CPU probe (test infrastructure, not collected by pytest): can the recurrent gradient states of the backward — a cell's dc
and its own dh_prev — live in fp16 at the loss scale S of dz?  Runs the oracle's explicit BPTT (oracle/convlstm_oracle.py
rollout_backward, source-patched in memory) and reports
  (1) per cell: max, 1 %-quantile and median of |S dz|, |S dh|, |S dc| over all steps (fp16: max 2^16, normal >= 2^-14);
  (2) the gradient error against the exact oracle with device-like fp16 operand rounding, with and without rounding
      dh / dc to fp16 (optionally times a power-of-two shift).
Results that shaped CLSTM_STATE16 (DESIGN.md finding 18): the three maxima agree within a bit; fp16 states add 1-6 % to the
worst gradient error (depth-36 x1: 8.55e-4 -> 9.06e-4, x3: 9.96e-4 -> 1.03e-3); a 2^-6 down-shift is harmful.
Usage: python tests/probe_state_ranges.py [ranges|errors]"""
import inspect, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import convlstm_oracle as O

SRC = inspect.getsource(O.rollout_backward)
LINE = "dh[k], dc[k] = dcomb[:, cx:], dc_prev"
assert LINE in SRC


def setup(B, tin, tout, cin, hid, cout, H, W, L=2, k=3, ws=1.0, seed=0):
    g = torch.Generator().manual_seed(1234 + seed)
    p = O.init_params(cin, hid, cout, n_layers=L, kernel_size=(k, k), seed=seed, cell_weight_scale=ws)
    x = torch.randn(B, tin, cin, H, W, generator=g)
    tgt = torch.rand(B, tout, cout, H, W, generator=g)
    y, sv = O.rollout_forward(x, p, tout, n_layers=L)
    loss, dy = O.mse_loss_and_grad(y, tgt)
    amax = (dy * sv.y * (1 - sv.y)).abs().max().item()
    S = 2.0 ** math.floor(math.log2(1024.0 / amax))  # the device's choice: max |S dlogit| in [2^9, 2^10]
    return p, x, tgt, sv, dy, S


def ranges(name, *a, **kw):
    p, x, tgt, sv, dy, S = setup(*a, **kw)
    stats = {}

    def rec(kc, dz, dh, dc):
        for nm, v in (("dz", dz), ("dh", dh), ("dc", dc)):
            v = v.abs().flatten()
            nz = v[v > 0]
            q = torch.quantile(nz[:: max(1, nz.numel() // 200000)], torch.tensor([0.01, 0.5]))
            e = stats.setdefault((kc, nm), [0.0, 1e30, 1e30])
            e[0], e[1], e[2] = max(e[0], v.max().item()), min(e[1], q[0].item()), min(e[2], q[1].item())

    ns = dict(O.__dict__)
    ns["REC"] = rec
    exec(SRC.replace(LINE, LINE + "; REC(k, dz * s, dh[k] * s, dc[k] * s)").replace("def rollout_backward", "def bwd"), ns)
    ns["bwd"](dy, sv, p, O.Rounding(dz_scale=S))
    print(name)
    for (kc, nm), (mx, q01, q50) in sorted(stats.items()):
        print(f"  cell {kc} {nm}: max 2^{math.log2(mx):6.2f}   smallest 1%-quantile 2^{math.log2(q01):7.2f}   smallest median 2^{math.log2(q50):7.2f}")


def errors(name, *a, shift=1.0, **kw):
    p, x, tgt, sv, dy, S = setup(*a, **kw)
    g_exact = O.rollout_backward(dy, sv, p)
    r = O.Rounding(act="fp16", weight="fp16", dz="fp16", gates="fp16", dz_scale=S)
    tout = tgt.shape[1]
    y_r, sv_r = O.rollout_forward(x, p, tout, n_layers=sv.n_layers, r=r)
    _, dy_r = O.mse_loss_and_grad(y_r, tgt)
    for dhk, dck in ((None, None), ("fp16", None), ("fp16", "fp16")):
        ns = dict(O.__dict__)
        ns.update(O=O, DH=dhk, DC=dck, SH=shift)
        exec(SRC.replace(LINE, "dh[k], dc[k] = O.round_to(dcomb[:, cx:] * s * SH, DH) / (s * SH), O.round_to(dc_prev * s * SH, DC) / (s * SH)")
             .replace("def rollout_backward", "def bwd"), ns)
        g_r = ns["bwd"](dy_r, sv_r, p, r)
        worst = sorted(((O.rel_l2(g_r[k], g_exact[k]), k) for k in g_exact), reverse=True)[:3]
        print(f"{name} shift {shift} dh={dhk} dc={dck}: " + " ".join(f"{k}={v:.2e}" for v, k in worst), flush=True)


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 1)
    what = sys.argv[1] if len(sys.argv) > 1 else "ranges"
    cases = (("depth-36 64px x1", (2, 12, 24, 12, 64, 12, 64, 64), {}),
             ("depth-36 64px x3", (2, 12, 24, 12, 64, 12, 64, 64), {"ws": 3.0}),
             ("3-layer k5 17x64", (1, 3, 2, 17, 64, 3, 17, 64), {"L": 3, "k": 5, "seed": 3}),
             ("x8 weights 6/8 32px", (2, 6, 8, 12, 64, 12, 32, 32), {"ws": 8.0}))
    for name, a, kw in cases:
        if what == "ranges":
            ranges(name, *a, **kw)
        else:
            errors(name, *a, **kw)
            if name.startswith("3-layer"):
                errors(name, *a, shift=1.0 / 64.0, **kw)
