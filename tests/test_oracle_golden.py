"""CPU: the oracle restatement against the golden fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py).  These run everywhere (no /root/reference, no GPU needed)."""
import numpy as np
import pytest
import torch

from oracle import convlstm_oracle as O
from gpu_checks import load_golden


@pytest.mark.parametrize("name", ["rollout_h16_12x10", "rollout_h8_stress"])
def test_rollout_oracle_matches_golden(name):
    z = load_golden(name)
    p = {k[len("param.model."):]: v for k, v in z.items() if k.startswith("param.model.")}
    tout = z["target"].shape[1]
    y, sv = O.rollout_forward(z["x"], p, tout)
    assert torch.allclose(y, z["y"], rtol=0, atol=1e-6)
    loss, dy = O.mse_loss_and_grad(y, z["target"])
    assert abs(loss.item() - float(z["loss"])) < 1e-7
    g = O.rollout_backward(dy, sv, p)
    for k, v in g.items():
        assert O.rel_l2(v, z["grad.model." + k]) < 2e-5, k


@pytest.mark.parametrize("name", ["cell_k3", "cell_k35"])
def test_cell_oracle_matches_golden(name):
    z = load_golden(name)
    hn, cn, gates = O.cell_forward(z["x"], z["h"], z["c"], z["weight"], z["bias"])
    assert torch.allclose(hn, z["h_next"], atol=1e-6) and torch.allclose(cn, z["c_next"], atol=1e-6)
    # explicit backward stages == reference autograd
    dz, dc_prev = O.cell_gate_grad(z["dh"], z["dc"], gates, z["c"], cn)
    cin = z["x"].shape[1]
    dcomb = O.conv_dgrad(dz, z["weight"])
    assert O.rel_l2(dcomb[:, :cin], z["dx"]) < 2e-5
    assert O.rel_l2(dcomb[:, cin:], z["dh_cur"]) < 2e-5
    assert O.rel_l2(dc_prev, z["dc_cur"]) < 2e-5
    dw = O.conv_wgrad(torch.cat([z["x"], z["h"]], 1), dz, z["weight"].shape[2:])
    assert O.rel_l2(dw, z["dweight"]) < 2e-5
    assert O.rel_l2(dz.sum(dim=(0, 2, 3)), z["dbias"]) < 2e-5


def test_low_precision_emulation_orders():
    """fp16 operands (10-bit mantissa) must beat bf16 on the gradient bar (DESIGN.md 'Numerics')."""
    torch.manual_seed(0)
    p = O.init_params(12, 16, 12, seed=0, cell_weight_scale=3.0)
    x = torch.randn(2, 4, 12, 16, 16)
    tgt = torch.rand(2, 5, 12, 16, 16)
    y, sv = O.rollout_forward(x, p, 5)
    _, dy = O.mse_loss_and_grad(y, tgt)
    g = O.rollout_backward(dy, sv, p)
    worst = {}
    for kind in ("fp16", "bf16"):
        r = O.Rounding(act=kind, weight=kind, dz=kind, gates=kind, dz_scale=2.0 ** 20 if kind == "fp16" else 1.0)
        yq, svq = O.rollout_forward(x, p, 5, r=r)
        _, dyq = O.mse_loss_and_grad(yq, tgt)
        gq = O.rollout_backward(dyq, svq, p, r=r)
        worst[kind] = max(O.rel_l2(gq[k], g[k]) for k in g)
    assert worst["fp16"] < 2e-3 < worst["bf16"] * 4
    assert worst["fp16"] < worst["bf16"]


def test_forecast_steps_zero_raises_like_reference():
    p = O.init_params(12, 8, 1)
    with pytest.raises(RuntimeError):
        O.rollout_forward(torch.randn(1, 2, 12, 8, 8), p, 0)
