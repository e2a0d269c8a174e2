"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.npz from the UNMODIFIED reference classes.

Run in the build container (needs /root/reference):  python -m oracle.make_golden
Each fixture holds the seeded inputs, the reference state_dict and what the reference computed:
forward output, MSE loss and every parameter gradient (torch CPU fp32, autograd).
"""
from __future__ import annotations

import os

import numpy as np
import torch

from oracle import convlstm_oracle as O
from oracle.reference_loader import load_reference

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    # name: (B, T_in, T_out, C_in, hid, C_out, H, W, weight_scale)
    "rollout_h16_12x10": (2, 3, 4, 12, 16, 5, 12, 10, 1.0),
    "rollout_h8_stress": (1, 2, 3, 12, 8, 12, 9, 16, 3.0),
}
CELL_CASES = {
    # name: (B, C_in, hid, H, W, kh, kw)
    "cell_k3": (2, 12, 8, 7, 9, 3, 3),
    "cell_k35": (1, 5, 4, 6, 8, 3, 5),
}


def main() -> None:
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    Cell, Net, Lit = load_reference()
    torch.set_num_threads(1)
    for name, (B, tin, tout, cin, hid, cout, H, W, ws) in CASES.items():
        g = torch.Generator().manual_seed(1234)
        p = O.init_params(cin, hid, cout, seed=0, cell_weight_scale=ws)
        x = torch.randn(B, tin, cin, H, W, generator=g)
        tgt = torch.rand(B, tout, cout, H, W, generator=g)
        lit = Lit(hidden_dim=hid, input_channels=cin, out_channels=cout, forecast_steps=tout)
        lit.load_state_dict({f"model.{k}": v for k, v in p.items()})
        loss = lit.training_step((x, tgt), 0)  # the reference's own training_step (conv_lstm.py:53-70)
        loss.backward()
        with torch.no_grad():
            y = lit(x, tout)
        out = {"x": x.numpy(), "target": tgt.numpy(), "y": y.numpy(), "loss": np.float32(loss.item())}
        for k, v in lit.named_parameters():
            out["param." + k] = v.detach().numpy()
            out["grad." + k] = v.grad.numpy()
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **out)
        print(name, "loss", loss.item(), {k: v.shape for k, v in out.items() if k.startswith("grad.")})
    # A checkpoint in the format Lightning's ModelCheckpoint(save_weights_only=True) writes for this module
    # (configs/callbacks/default.yaml:1-17): `state_dict` + the `hyper_parameters` of save_hyperparameters()
    # (conv_lstm.py:33), produced by the reference class itself, plus one seeded input and the reference's output for it.
    torch.manual_seed(11)
    hp = dict(hidden_dim=16, input_channels=12, out_channels=3, forecast_steps=5, lr=1e-4, visualize=False, loss="mse",
              pretrained=False, conv_type="standard")
    lit = Lit(**hp)
    g = torch.Generator().manual_seed(99)
    x = torch.randn(2, 3, 12, 10, 12, generator=g)
    with torch.no_grad():
        y = lit(x, hp["forecast_steps"])
    torch.save({"epoch": 3, "global_step": 1234, "pytorch-lightning_version": "1.4.9",
                "state_dict": {k: v.clone() for k, v in lit.state_dict().items()}, "hyper_parameters": hp,
                "golden": {"x": x, "y": y}}, os.path.join(GOLDEN_DIR, "lightning_best.ckpt"))
    print("lightning_best.ckpt ok")
    for name, (B, cin, hid, H, W, kh, kw) in CELL_CASES.items():
        torch.manual_seed(7)
        cell = Cell(cin, hid, (kh, kw), True)
        x = torch.randn(B, cin, H, W, requires_grad=True)
        h = torch.randn(B, hid, H, W, requires_grad=True)
        c = torch.randn(B, hid, H, W, requires_grad=True)
        dh = torch.randn(B, hid, H, W)
        dc = torch.randn(B, hid, H, W)
        hn, cn = cell(x, [h, c])
        (hn * dh + cn * dc).sum().backward()
        out = {
            "x": x.detach().numpy(), "h": h.detach().numpy(), "c": c.detach().numpy(), "dh": dh.numpy(),
            "dc": dc.numpy(), "weight": cell.conv.weight.detach().numpy(), "bias": cell.conv.bias.detach().numpy(),
            "h_next": hn.detach().numpy(), "c_next": cn.detach().numpy(), "dx": x.grad.numpy(),
            "dh_cur": h.grad.numpy(), "dc_cur": c.grad.numpy(), "dweight": cell.conv.weight.grad.numpy(),
            "dbias": cell.conv.bias.grad.numpy(),
        }
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **out)
        print(name, "ok")


if __name__ == "__main__":
    main()
