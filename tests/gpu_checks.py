"""Shared parity checks: CUDA path (through the C ABI) vs the CPU fp32 oracle / golden fixtures.
Each returns {tensor name: rel-L2}; the tests assert the tolerance, tools/gpu_check.py prints them."""
from __future__ import annotations

import os
from typing import Dict

import numpy as np
import torch

from oracle import convlstm_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 2e-3  # BASELINE.json north_star: relative-L2 <= 2e-3 per tensor


def rel(a: torch.Tensor, b: torch.Tensor) -> float:
    return O.rel_l2(a.detach().cpu().float(), b.detach().cpu().float())


def load_golden(name: str) -> Dict[str, torch.Tensor]:
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: torch.from_numpy(np.asarray(z[k])) for k in z.files}


def cell_case(B, cin, hid, H, W, kh, kw, dtype="fp16", seed=0, backward=True) -> Dict[str, float]:
    from satflow_b200 import ConvLSTMCell

    g = torch.Generator().manual_seed(seed)
    # the cell's weights come from torch's default init, i.e. from the GLOBAL generator: seed it, otherwise the case
    # depends on whatever ran before it (a 3-pixel case then misses the tolerance once in ~40 runs by rounding luck)
    torch.manual_seed(1000 + seed)
    cell = ConvLSTMCell(cin, hid, (kh, kw), True)
    cell.operand_dtype = dtype
    x = torch.randn(B, cin, H, W, generator=g)
    h = torch.randn(B, hid, H, W, generator=g) * 0.5
    c = torch.randn(B, hid, H, W, generator=g)
    dh = torch.randn(B, hid, H, W, generator=g)
    dc = torch.randn(B, hid, H, W, generator=g)
    w = cell.conv.weight.detach().clone()
    b = cell.conv.bias.detach().clone()
    # oracle
    xo, ho, co = (t.clone().requires_grad_(True) for t in (x, h, c))
    wo, bo = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    hn_o, cn_o, _ = O.cell_forward(xo, ho, co, wo, bo)
    out: Dict[str, float] = {}
    cell = cell.cuda()
    xg, hg, cg = (t.cuda().requires_grad_(True) for t in (x, h, c))
    hn, cn = cell(xg, [hg, cg])
    out["h_next"] = rel(hn, hn_o)
    out["c_next"] = rel(cn, cn_o)
    if backward:
        (hn_o * dh + cn_o * dc).sum().backward()
        (hn * dh.cuda() + cn * dc.cuda()).sum().backward()
        out["dx"] = rel(xg.grad, xo.grad)
        out["dh_cur"] = rel(hg.grad, ho.grad)
        out["dc_cur"] = rel(cg.grad, co.grad)
        out["dweight"] = rel(cell.conv.weight.grad, wo.grad)
        out["dbias"] = rel(cell.conv.bias.grad, bo.grad)
    return out


def cell_golden(name: str, dtype="fp16") -> Dict[str, float]:
    from satflow_b200 import ConvLSTMCell

    z = load_golden(name)
    hid = z["h"].shape[1]
    cin = z["x"].shape[1]
    kh, kw = z["weight"].shape[2:]
    cell = ConvLSTMCell(cin, hid, (kh, kw), True)
    cell.operand_dtype = dtype
    cell.load_state_dict({"conv.weight": z["weight"], "conv.bias": z["bias"]})
    cell = cell.cuda()
    xg, hg, cg = (z[k].cuda().requires_grad_(True) for k in ("x", "h", "c"))
    hn, cn = cell(xg, [hg, cg])
    (hn * z["dh"].cuda() + cn * z["dc"].cuda()).sum().backward()
    return {
        "h_next": rel(hn, z["h_next"]), "c_next": rel(cn, z["c_next"]), "dx": rel(xg.grad, z["dx"]),
        "dh_cur": rel(hg.grad, z["dh_cur"]), "dc_cur": rel(cg.grad, z["dc_cur"]),
        "dweight": rel(cell.conv.weight.grad, z["dweight"]), "dbias": rel(cell.conv.bias.grad, z["dbias"]),
    }


def rollout_case(B, tin, tout, cin, hid, cout, H, W, n_layers=2, k=3, dtype="fp16", weight_scale=1.0, seed=0,
                 backward=True, states=True, loss_scale=1.0, status=None) -> Dict[str, float]:
    """CUDA rollout (module API -> C ABI) vs the oracle on the same seeded inputs and weights.  ``loss_scale``
    multiplies the loss on both sides (gradient range tests); ``status`` (a dict) receives the range statistics of
    the device backward."""
    from satflow_b200 import ConvLSTM

    g = torch.Generator().manual_seed(1234 + seed)
    p = O.init_params(cin, hid, cout, n_layers=n_layers, kernel_size=(k, k), seed=seed, cell_weight_scale=weight_scale)
    x = torch.randn(B, tin, cin, H, W, generator=g)
    tgt = torch.rand(B, tout, cout, H, W, generator=g)
    y_o, sv = O.rollout_forward(x, p, tout, n_layers=n_layers)
    net = ConvLSTM(cin, hid, cout, n_layers=n_layers, kernel_size=(k, k), operand_dtype=dtype)
    net.load_state_dict(p)
    net = net.cuda()
    out: Dict[str, float] = {}
    if backward:
        y = net(x.cuda(), tout)
        loss = torch.nn.functional.mse_loss(y.permute(0, 2, 1, 3, 4), tgt.cuda())
        (loss * loss_scale).backward()
        st = net.check_gradients()  # raises FloatingPointError on a 16-bit overflow
        if status is not None and st is not None:
            status.update(st)
        loss_o, dy_o = O.mse_loss_and_grad(y_o, tgt)
        g_o = O.rollout_backward(dy_o * loss_scale, sv, p)
        out["loss"] = abs(loss.item() - loss_o.item()) / abs(loss_o.item())
        for name, prm in net.named_parameters():
            out["grad." + name] = rel(prm.grad, g_o[name])
    else:
        with torch.no_grad():
            y = net(x.cuda(), tout)
    out["y"] = rel(y, y_o)
    # pre-sigmoid logits are the sensitive forward signal
    yc = y.detach().cpu().double().clamp(1e-12, 1 - 1e-12)
    out["logits"] = rel(torch.log(yc / (1 - yc)).float(), sv.logits)
    if states:
        plan = next(iter(net._plans.values()))
        ncell = 2 * n_layers
        for cidx in range(ncell):
            T = tin if cidx < n_layers else tout
            h, c = plan.read_state(cidx, T)
            out[f"h_final[{cidx}]"] = rel(h, sv.final_h[cidx])
            out[f"c_final[{cidx}]"] = rel(c, sv.final_c[cidx])
    net.release_plans()
    return out


def rollout_golden(name: str, dtype="fp16") -> Dict[str, float]:
    from satflow_b200 import EncoderDecoderConvLSTM

    z = load_golden(name)
    x, tgt = z["x"], z["target"]
    hid = z["param.model.encoder_1_convlstm.conv.bias"].numel() // 4
    cout = z["param.model.decoder_CNN.bias"].numel()
    lit = EncoderDecoderConvLSTM(hidden_dim=hid, input_channels=x.shape[2], out_channels=cout, forecast_steps=tgt.shape[1])
    lit.model.operand_dtype = dtype
    lit.load_state_dict({k[len("param."):]: v for k, v in z.items() if k.startswith("param.")})
    lit = lit.cuda()
    loss = lit.training_step((x.cuda(), tgt.cuda()), 0)
    loss.backward()
    with torch.no_grad():
        y = lit(x.cuda(), tgt.shape[1])
    out = {"y": rel(y, z["y"]), "loss": abs(loss.item() - float(z["loss"])) / float(z["loss"])}
    for k, prm in lit.named_parameters():
        out["grad." + k] = rel(prm.grad, z["grad." + k])
    lit.model.release_plans()
    return out


def cell_unrolled_case(B, cin, hid, H, W, steps, k=3, seed=0) -> Dict[str, float]:
    """The reference's ONLY way of using the cell (conv_lstm.py:176-183): the same ConvLSTMCell applied ``steps``
    times, h/c threaded through, then loss.backward() through all of them.  Every step's gradient contributes to the
    shared weight, so each autograd node must own its saved activations."""
    from satflow_b200 import ConvLSTMCell

    g = torch.Generator().manual_seed(77 + seed)
    torch.manual_seed(2000 + seed)
    cell = ConvLSTMCell(cin, hid, (k, k), True)
    xs = torch.randn(steps, B, cin, H, W, generator=g)
    wsum = torch.randn(steps, B, hid, H, W, generator=g)  # loss = sum_t <h_t, wsum_t> + <c_T, wsum_0>
    w = cell.conv.weight.detach().clone()
    b = cell.conv.bias.detach().clone()
    # oracle (autograd through the restated cell)
    wo, bo = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    xo = xs.clone().requires_grad_(True)
    h = torch.zeros(B, hid, H, W)
    c = torch.zeros(B, hid, H, W)
    loss_o = 0.0
    for t in range(steps):
        h, c, _ = O.cell_forward(xo[t], h, c, wo, bo)
        loss_o = loss_o + (h * wsum[t]).sum()
    loss_o = loss_o + (c * wsum[0]).sum()
    loss_o.backward()
    # device
    cell = cell.cuda()
    xg = xs.cuda().requires_grad_(True)
    hg, cg = cell.init_hidden(B, (H, W))
    loss = 0.0
    wg = wsum.cuda()
    for t in range(steps):
        hg, cg = cell(xg[t], [hg, cg])
        loss = loss + (hg * wg[t]).sum()
    loss = loss + (cg * wg[0]).sum()
    loss.backward()
    # (the scalar loss is a sum of randomly signed terms, dominated by cancellation: not a parity signal)
    return {
        "h_T": rel(hg, h), "c_T": rel(cg, c), "dx": rel(xg.grad, xo.grad),
        "dweight": rel(cell.conv.weight.grad, wo.grad), "dbias": rel(cell.conv.bias.grad, bo.grad),
    }


def cloudgan_generator_case(B=2, tin=3, tout=4, cin=12, hid=16, cout=12, H=16, W=24) -> Dict[str, float]:
    """What CloudGAN does to a ConvLSTM generator (cloudgan.py:88-92, gan/generators.py:49-50,69, gan/common.py:44-64):
    build it with keyword arguments, take a generator step through it, RE-INITIALISE every Conv weight through
    ``m.weight.data`` (which does not bump the autograd version counter), run it again and slice the (B, C, T, H, W)
    output per time step (cloudgan.py:147,176,291); then clip the weights in place (WGAN) and run once more.  Every
    forward must see the weights of that moment."""
    from torch.nn import init

    from satflow_b200 import ConvLSTM

    def states(net_, sv_, out_, tag):
        plan = [p_ for p_ in net_._plans.values() if not p_.training][-1]
        for cidx in range(4):
            h_, c_ = plan.read_state(cidx, tin if cidx < 2 else tout)
            out_[f"{tag}.h_final[{cidx}]"] = rel(h_, sv_.final_h[cidx])
            out_[f"{tag}.c_final[{cidx}]"] = rel(c_, sv_.final_c[cidx])

    torch.manual_seed(31)
    net = ConvLSTM(cin, hidden_dim=hid, out_channels=cout).cuda()
    g = torch.Generator().manual_seed(32)
    x = torch.randn(B, tin, cin, H, W, generator=g)
    tgt = torch.rand(B, tout, cout, H, W, generator=g)
    out: Dict[str, float] = {}
    # a generator step on the freshly constructed net (torch default init): forward, loss, backward
    p0 = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    y0_o, sv0 = O.rollout_forward(x, p0, tout)
    yg = net(x.cuda(), forecast_steps=tout)
    torch.nn.functional.mse_loss(yg.permute(0, 2, 1, 3, 4), tgt.cuda()).backward()
    _, dy_o = O.mse_loss_and_grad(y0_o, tgt)
    g_o = O.rollout_backward(dy_o, sv0, p0)
    out["y_before"] = rel(yg, y0_o)
    for name, prm in net.named_parameters():
        out["grad." + name] = rel(prm.grad, g_o[name])
    y_before = yg.detach().clone()

    def init_func(m):  # gan/common.py:44-64, init_type "normal", gain 0.02
        classname = m.__class__.__name__
        if hasattr(m, "weight") and (classname.find("Conv") != -1 or classname.find("Linear") != -1):
            init.normal_(m.weight.data, 0.0, 0.02)
            if hasattr(m, "bias") and m.bias is not None:
                init.constant_(m.bias.data, 0.0)

    net.apply(init_func)
    p = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    y_o, sv = O.rollout_forward(x, p, tout)
    with torch.no_grad():
        y = net(x.cuda(), forecast_steps=tout)
    out["y_after_reinit"] = rel(y, y_o)
    states(net, sv, out, "reinit")  # y is ~0.5 everywhere after N(0, 0.02): the states are the sensitive signal
    out["changed"] = 0.0 if float((y - y_before).abs().max()) > 1e-4 else 1.0  # 1.0 = stale weights were used
    for i in range(tout):  # consumers slice generated_images[:, :, i, :, :]
        out[f"slice[{i}]"] = rel(y[:, :, i, :, :], y_o[:, :, i, :, :])
    for prm in net.parameters():  # WGAN-style in-place clipping through .data
        prm.data.clamp_(-0.01, 0.01)
    p2 = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    y2_o, sv2 = O.rollout_forward(x, p2, tout)
    with torch.no_grad():
        out["y_after_clip"] = rel(net(x.cuda(), forecast_steps=tout), y2_o)
    states(net, sv2, out, "clip")
    net.release_plans()
    return out
