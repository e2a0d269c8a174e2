"""CPU probe (test infrastructure, not collected): where does the ~1e-3 gradient error of the 16-bit path come from?
The oracle with device-like fp16 rounding, one rounding class switched off at a time (depth-36 probe, default weights):
conv operands x / h (`act`), packed weights (`weight`), gate-gradient operands (`dz`), saved gates (`gates`)."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import convlstm_oracle as O

torch.set_num_threads(os.cpu_count() or 1)
ws = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
g = torch.Generator().manual_seed(1234)
p = O.init_params(12, 64, 12, seed=0, cell_weight_scale=ws)
x = torch.randn(2, 12, 12, 64, 64, generator=g)
tgt = torch.rand(2, 24, 12, 64, 64, generator=g)
y, sv = O.rollout_forward(x, p, 24)
loss, dy = O.mse_loss_and_grad(y, tgt)
g_exact = O.rollout_backward(dy, sv, p)
amax = (dy * sv.y * (1 - sv.y)).abs().max().item()
S = 2.0 ** math.floor(math.log2(1024.0 / amax))
full = dict(act="fp16", weight="fp16", dz="fp16", gates="fp16")
for off in (None, "act", "weight", "dz", "gates"):
    kw = dict(full)
    if off:
        kw[off] = None
    r = O.Rounding(dz_scale=S, **kw)
    y_r, sv_r = O.rollout_forward(x, p, 24, r=r)
    _, dy_r = O.mse_loss_and_grad(y_r, tgt)
    g_r = O.rollout_backward(dy_r, sv_r, p, r)
    worst = sorted(((O.rel_l2(g_r[k], g_exact[k]), k) for k in g_exact), reverse=True)[:3]
    print(f"x{ws} exact {off or '-'}: " + " ".join(f"{k}={v:.2e}" for v, k in worst) +
          f" | h_final[1] {O.rel_l2(sv_r.final_h[1], sv.final_h[1]):.1e}", flush=True)
