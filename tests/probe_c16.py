"""CPU probe (test infrastructure, not collected): what does storing the cell state c in fp16 between steps cost?  The oracle
forward with device-like fp16 operand rounding, with and without rounding the STORED c (h is computed from the fp32 c).
Result (DESIGN.md finding 19): final states 3e-4 -> 4-5.5e-4, worst gradient +1.5 %."""
import sys, inspect, math, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import convlstm_oracle as O
torch.set_num_threads(8)
src = inspect.getsource(O.cell_forward)
assert "h_next = o * torch.tanh(c_next)  # :55" in src
src2 = src.replace("h_next = o * torch.tanh(c_next)  # :55", "h_next = o * torch.tanh(c_next)\n    c_next = round_to(c_next, CKIND)")
orig = O.cell_forward
def patched(kind):
    O.__dict__["CKIND"] = kind
    exec(src2, O.__dict__)   # redefines O.cell_forward
for name, a, kw in (("depth-36 x1", (2, 12, 24, 12, 64, 12, 64, 64), {}), ("depth-36 x3", (2, 12, 24, 12, 64, 12, 64, 64), {"ws": 3.0})):
    B, tin, tout, cin, hid, cout, H, W = a
    g = torch.Generator().manual_seed(1234)
    p = O.init_params(cin, hid, cout, seed=0, cell_weight_scale=kw.get("ws", 1.0))
    x = torch.randn(B, tin, cin, H, W, generator=g)
    tgt = torch.rand(B, tout, cout, H, W, generator=g)
    O.cell_forward = orig
    y, sv = O.rollout_forward(x, p, tout)
    loss, dy = O.mse_loss_and_grad(y, tgt)
    g_exact = O.rollout_backward(dy, sv, p)
    amax = (dy * sv.y * (1 - sv.y)).abs().max().item()
    S = 2.0 ** math.floor(math.log2(1024.0 / amax))
    r = O.Rounding(act="fp16", weight="fp16", dz="fp16", gates="fp16", dz_scale=S)
    for ck in (None, "fp16"):
        patched(ck)
        y_r, sv_r = O.rollout_forward(x, p, tout, r=r)
        _, dy_r = O.mse_loss_and_grad(y_r, tgt)
        g_r = O.rollout_backward(dy_r, sv_r, p, r)
        worst = sorted(((O.rel_l2(g_r[k], g_exact[k]), k) for k in g_exact), reverse=True)[:3]
        st = " ".join(f"h{c}={O.rel_l2(sv_r.final_h[c], sv.final_h[c]):.1e}/c{c}={O.rel_l2(sv_r.final_c[c], sv.final_c[c]):.1e}" for c in range(4))
        print(f"{name} c={ck}: " + " ".join(f"{k}={v:.2e}" for v, k in worst) + " | logits " + f"{O.rel_l2(sv_r.logits, sv.logits):.1e} | " + st, flush=True)
