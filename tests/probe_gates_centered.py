"""CPU probe (test infrastructure, not collected): sigmoid gates saved as (gate - 0.5) in fp16 instead of the gate itself.
Oracle with device-like rounding, depth-36 probe: worst gradient 8.55e-4 -> 6.16e-4 (x3 weights: 9.96e-4 -> 8.02e-4), i.e. the
whole gate-rounding term of tests/probe_error_budget.py disappears (DESIGN.md finding 20)."""
import sys, math, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import convlstm_oracle as O
torch.set_num_threads(8)
orig = O.round_to
def round_to(t, kind):
    if kind == "fp16c":   # sigmoid gates stored as (gate - 0.5); tanh gate (can be negative) as is
        return torch.where(t.min() >= 0, (t - 0.5).to(torch.float16).to(torch.float32) + 0.5, t.to(torch.float16).to(torch.float32))
    return orig(t, kind)
O.round_to = round_to
for ws in (1.0, 3.0):
    g = torch.Generator().manual_seed(1234)
    p = O.init_params(12, 64, 12, seed=0, cell_weight_scale=ws)
    x = torch.randn(2, 12, 12, 64, 64, generator=g)
    tgt = torch.rand(2, 24, 12, 64, 64, generator=g)
    y, sv = O.rollout_forward(x, p, 24)
    loss, dy = O.mse_loss_and_grad(y, tgt)
    g_exact = O.rollout_backward(dy, sv, p)
    amax = (dy * sv.y * (1 - sv.y)).abs().max().item()
    S = 2.0 ** math.floor(math.log2(1024.0 / amax))
    for gk in ("fp16", "fp16c"):
        r = O.Rounding(act="fp16", weight="fp16", dz="fp16", gates=gk, dz_scale=S)
        y_r, sv_r = O.rollout_forward(x, p, 24, r=r)
        _, dy_r = O.mse_loss_and_grad(y_r, tgt)
        g_r = O.rollout_backward(dy_r, sv_r, p, r)
        worst = sorted(((O.rel_l2(g_r[k], g_exact[k]), k) for k in g_exact), reverse=True)[:2]
        print(f"x{ws} gates={gk}: " + " ".join(f"{k}={v:.2e}" for v, k in worst), flush=True)
