"""TEST INFRASTRUCTURE ONLY — CPU fp32 restatement of the reference ConvLSTM hot path.

Restates, in plain functional torch (CPU, fp32), the algorithm of

* ``ConvLSTMCell.forward``      /root/reference/satflow/models/layers/ConvLSTM.py:42-57
* ``ConvLSTMCell.init_hidden``  /root/reference/satflow/models/layers/ConvLSTM.py:59-64
* ``ConvLSTM.autoencoder``      /root/reference/satflow/models/conv_lstm.py:171-203
* ``ConvLSTM.forward``          /root/reference/satflow/models/conv_lstm.py:205-228
* ``EncoderDecoderConvLSTM.training_step`` (MSE loss) conv_lstm.py:53-70

and the backward pass those lines imply (SURVEY.md §8(a) "Backward math"), written out
explicitly as the same stages the CUDA path uses (gate-gradient pointwise, dgrad, wgrad)
so each device kernel has a one-to-one CPU counterpart.

Parity pinning: the reference's own tests hold NO golden vector for this path
(tests/test_models.py only constructs models), so this oracle is pinned against the
UNMODIFIED reference classes executed in the build container
(``oracle/reference_loader.py``; see ``tests/test_oracle_vs_reference.py``) and against
the fixtures those classes produced (``tests/golden/*.npz`` made by ``oracle/make_golden.py``).

Generalisation beyond the reference: ``n_layers`` (reference: fixed 2 encoder + 2 decoder
cells) and the kernel size (reference ConvLSTM: fixed 3x3) follow exactly the loop pattern
of conv_lstm.py:171-203 (L encoder cells chained, L decoder cells chained with
last->first feedback, zero initial states, same Conv3d(1,3,3)+Sigmoid head); for
n_layers=2, kernel 3 it is checked bit-for-bit against the reference.

An optional ``Rounding`` hook emulates low-precision operand storage (bf16 / fp16 / tf32
with fp32 accumulation) at exactly the points where the CUDA path rounds, which is how the
operand precision of the device kernels was chosen (DESIGN.md "Numerics").
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------
# parameter helpers (state_dict layout of conv_lstm.py:122-169)
# --------------------------------------------------------------------------------------
def cell_names(n_layers: int = 2) -> List[str]:
    return [f"encoder_{l + 1}_convlstm" for l in range(n_layers)] + [
        f"decoder_{l + 1}_convlstm" for l in range(n_layers)
    ]


def init_params(
    input_channels: int,
    hidden_dim: int,
    out_channels: int,
    n_layers: int = 2,
    kernel_size: Tuple[int, int] = (3, 3),
    seed: int = 0,
    cell_weight_scale: float = 1.0,
) -> Dict[str, Tensor]:
    """Random parameters with torch's Conv2d/Conv3d default init (what the reference gets
    from nn.Conv2d at layers/ConvLSTM.py:34-40 and nn.Conv3d at conv_lstm.py:164-169),
    keyed like the reference state_dict (without the ``model.`` prefix)."""
    g = torch.Generator().manual_seed(seed)
    p: Dict[str, Tensor] = {}
    kh, kw = kernel_size

    def uniform(shape, bound):
        return (torch.rand(shape, generator=g) * 2 - 1) * bound

    for idx, name in enumerate(cell_names(n_layers)):
        cin = (input_channels if idx == 0 else hidden_dim) + hidden_dim
        bound = 1.0 / (cin * kh * kw) ** 0.5
        p[f"{name}.conv.weight"] = uniform((4 * hidden_dim, cin, kh, kw), bound) * cell_weight_scale
        p[f"{name}.conv.bias"] = uniform((4 * hidden_dim,), bound)
    bound = 1.0 / (hidden_dim * 9) ** 0.5
    p["decoder_CNN.weight"] = uniform((out_channels, hidden_dim, 1, 3, 3), bound)
    p["decoder_CNN.bias"] = uniform((out_channels,), bound)
    return p


# --------------------------------------------------------------------------------------
# low-precision emulation hook
# --------------------------------------------------------------------------------------
def round_to(t: Tensor, kind: Optional[str]) -> Tensor:
    if kind is None or kind == "fp32":
        return t
    if kind == "bf16":
        return t.to(torch.bfloat16).to(torch.float32)
    if kind == "fp16":
        return t.to(torch.float16).to(torch.float32)
    if kind == "tf32":  # 10-bit mantissa, round-to-nearest-even on the raw bits
        i = t.contiguous().view(torch.int32)
        lsb = (i >> 13) & 1
        i = (i + 0xFFF + lsb) & ~0x1FFF
        return i.view(torch.float32)
    raise ValueError(kind)


@dataclass
class Rounding:
    """Where the device path rounds.  None everywhere == the exact fp32 oracle."""

    act: Optional[str] = None  # conv operand storage of x and h (fwd A operand, wgrad B operand)
    weight: Optional[str] = None  # packed conv weights (fwd / dgrad B operand)
    dz: Optional[str] = None  # gate pre-activation gradients (dgrad A / wgrad A operand)
    gates: Optional[str] = None  # saved post-activation gates i,f,o,g
    dz_scale: float = 1.0  # static power-of-two scale applied before rounding dz
    head_act: Optional[str] = None  # h-stack operand of the head conv (defaults to act)

    def q(self, t: Tensor, what: str) -> Tensor:
        kind = getattr(self, what)
        if what == "head_act" and kind is None:
            kind = self.act
        return round_to(t, kind)


EXACT = Rounding()


# --------------------------------------------------------------------------------------
# one cell step, forward and backward
# --------------------------------------------------------------------------------------
def cell_forward(
    x: Tensor, h: Tensor, c: Tensor, weight: Tensor, bias: Optional[Tensor], r: Rounding = EXACT
):
    """layers/ConvLSTM.py:42-57.  Gate order along dim 0 of ``weight`` is i, f, o, g (:48)."""
    hid = h.shape[1]
    kh, kw = weight.shape[2], weight.shape[3]
    combined = torch.cat([r.q(x, "act"), r.q(h, "act")], dim=1)  # :45
    z = F.conv2d(combined, r.q(weight, "weight"), bias, padding=(kh // 2, kw // 2))  # :47
    zi, zf, zo, zg = torch.split(z, hid, dim=1)  # :48
    i, f, o, g = torch.sigmoid(zi), torch.sigmoid(zf), torch.sigmoid(zo), torch.tanh(zg)  # :49-52
    c_next = f * c + i * g  # :54
    h_next = o * torch.tanh(c_next)  # :55
    return h_next, c_next, (i, f, o, g)


def cell_gate_grad(dh: Tensor, dc: Tensor, gates, c_prev: Tensor, c_next: Tensor):
    """Pointwise backward of :49-55.  Returns dz (N order i,f,o,g) and dc_prev."""
    i, f, o, g = gates
    tc = torch.tanh(c_next)
    do = dh * tc
    dc_tot = dc + dh * o * (1 - tc * tc)
    dz = torch.cat(
        [
            dc_tot * g * i * (1 - i),
            dc_tot * c_prev * f * (1 - f),
            do * o * (1 - o),
            dc_tot * i * (1 - g * g),
        ],
        dim=1,
    )
    return dz, dc_tot * f


def conv_dgrad(dz: Tensor, weight: Tensor) -> Tensor:
    kh, kw = weight.shape[2], weight.shape[3]
    return F.conv_transpose2d(dz, weight, padding=(kh // 2, kw // 2))


def conv_wgrad(inp: Tensor, dz: Tensor, kernel_size) -> Tensor:
    """dW[n, c, dy, dx] = sum_{b,y,x} dz[b,n,y,x] * inp[b,c,y+dy-ph,x+dx-pw]."""
    kh, kw = kernel_size
    B, C, H, W = inp.shape
    cols = F.unfold(inp, (kh, kw), padding=(kh // 2, kw // 2))  # (B, C*kh*kw, H*W)
    dzf = dz.reshape(B, dz.shape[1], H * W)
    dw = torch.einsum("bnp,bkp->nk", dzf, cols)
    return dw.reshape(dz.shape[1], C, kh, kw)


# --------------------------------------------------------------------------------------
# rollout forward / backward (conv_lstm.py:171-228)
# --------------------------------------------------------------------------------------
@dataclass
class Saved:
    x: Tensor
    n_layers: int
    t_in: int
    t_out: int
    # per cell (encoder cells then decoder cells): per step lists
    inp: List[List[Tensor]] = field(default_factory=list)
    h_prev: List[List[Tensor]] = field(default_factory=list)
    c_prev: List[List[Tensor]] = field(default_factory=list)
    c_next: List[List[Tensor]] = field(default_factory=list)
    gates: List[List[tuple]] = field(default_factory=list)
    h_stack: Optional[Tensor] = None
    logits: Optional[Tensor] = None
    y: Optional[Tensor] = None
    final_h: Optional[List[Tensor]] = None
    final_c: Optional[List[Tensor]] = None


def rollout_forward(
    x: Tensor,
    params: Dict[str, Tensor],
    forecast_steps: int,
    n_layers: int = 2,
    r: Rounding = EXACT,
) -> Tuple[Tensor, Saved]:
    """ConvLSTM.forward: x (B,T_in,C,H,W) -> (B,C_out,T_out,H,W)."""
    if forecast_steps <= 0:
        # conv_lstm.py:198 -> torch.stack([]) raises RuntimeError in the reference
        raise RuntimeError("stack expects a non-empty TensorList")
    B, T_in, _, H, W = x.shape
    names = cell_names(n_layers)
    hid = params[f"{names[0]}.conv.bias"].numel() // 4
    ncell = 2 * n_layers
    zeros = lambda: torch.zeros(B, hid, H, W, dtype=torch.float32, device=x.device)  # layers/ConvLSTM.py:59-64
    h = [zeros() for _ in range(ncell)]  # :218-221
    c = [zeros() for _ in range(ncell)]
    sv = Saved(x=x, n_layers=n_layers, t_in=T_in, t_out=forecast_steps)
    for lst in (sv.inp, sv.h_prev, sv.c_prev, sv.c_next, sv.gates):
        lst.extend([[] for _ in range(ncell)])

    def step(k: int, inp: Tensor):
        w, b = params[f"{names[k]}.conv.weight"], params[f"{names[k]}.conv.bias"]
        hn, cn, gates = cell_forward(inp, h[k], c[k], w, b, r)
        sv.inp[k].append(inp)
        sv.h_prev[k].append(h[k])
        sv.c_prev[k].append(c[k])
        sv.c_next[k].append(cn)
        sv.gates[k].append(tuple(r.q(gt, "gates") for gt in gates))
        h[k], c[k] = hn, cn

    for t in range(T_in):  # :176-183
        step(0, x[:, t])
        for l in range(1, n_layers):
            step(l, h[l - 1])
    enc_vec = h[n_layers - 1]  # :185
    outs = []
    for t in range(forecast_steps):  # :188-196
        step(n_layers, enc_vec)
        for l in range(1, n_layers):
            step(n_layers + l, h[n_layers + l - 1])
        enc_vec = h[ncell - 1]
        outs.append(enc_vec)
    h_stack = torch.stack(outs, 1).permute(0, 2, 1, 3, 4)  # :198-199 (B,hid,T,H,W)
    logits = F.conv3d(
        r.q(h_stack, "head_act"),
        r.q(params["decoder_CNN.weight"], "weight"),
        params["decoder_CNN.bias"],
        padding=(0, 1, 1),
    )  # :200
    y = torch.sigmoid(logits)  # :201
    sv.h_stack, sv.logits, sv.y = h_stack, logits, y
    sv.final_h, sv.final_c = h, c
    return y, sv


def rollout_backward(
    dy: Tensor, sv: Saved, params: Dict[str, Tensor], r: Rounding = EXACT
) -> Dict[str, Tensor]:
    """Explicit BPTT through ``rollout_forward``.  ``dy`` is dL/dy with y (B,C_out,T_out,H,W).
    Returns parameter gradients keyed like ``params``."""
    L, T_in, T_out = sv.n_layers, sv.t_in, sv.t_out
    names = cell_names(L)
    ncell = 2 * L
    B, hid = sv.h_stack.shape[0], sv.h_stack.shape[1]
    H, W = sv.h_stack.shape[3], sv.h_stack.shape[4]
    grads = {k: torch.zeros_like(v) for k, v in params.items()}
    s = r.dz_scale

    # head: sigmoid + Conv3d(1,3,3)
    dlog = dy * sv.y * (1 - sv.y)
    dlog_q = r.q(dlog * s, "dz") / s
    w_head = params["decoder_CNN.weight"]
    grads["decoder_CNN.bias"] = dlog.sum(dim=(0, 2, 3, 4))
    dstack = torch.zeros_like(sv.h_stack)
    for t in range(T_out):
        hs = r.q(sv.h_stack[:, :, t], "head_act")
        grads["decoder_CNN.weight"][:, :, 0] += conv_wgrad(hs, dlog_q[:, :, t], (3, 3))
        dstack[:, :, t] = conv_dgrad(dlog_q[:, :, t], r.q(w_head[:, :, 0], "weight"))

    zeros = lambda: torch.zeros(B, hid, H, W)
    dh = [zeros() for _ in range(ncell)]  # gradient wrt each cell's current h (from its own next step)
    dc = [zeros() for _ in range(ncell)]

    def back_step(k: int, t: int, dh_extra: Tensor) -> Tensor:
        """Backward of cell k at its step index t.  dh_extra = grad arriving at h'_t from
        consumers other than the cell's own next step.  Returns dx (grad wrt the input)."""
        w = params[f"{names[k]}.conv.weight"]
        cx = w.shape[1] - hid
        dz, dc_prev = cell_gate_grad(
            dh[k] + dh_extra, dc[k], sv.gates[k][t], sv.c_prev[k][t], sv.c_next[k][t]
        )
        grads[f"{names[k]}.conv.bias"] += dz.sum(dim=(0, 2, 3))
        dz_q = r.q(dz * s, "dz") / s
        comb = torch.cat([r.q(sv.inp[k][t], "act"), r.q(sv.h_prev[k][t], "act")], 1)
        grads[f"{names[k]}.conv.weight"] += conv_wgrad(comb, dz_q, w.shape[2:])
        dcomb = conv_dgrad(dz_q, r.q(w, "weight"))
        dh[k], dc[k] = dcomb[:, cx:], dc_prev
        return dcomb[:, :cx]

    # decoder, reverse time.  dec-last h'_t feeds: head (dstack[t]), dec-first at t+1 (feedback :195)
    dfeed = zeros()  # grad wrt the decoder input of step t+1 (encoder_vector)
    for t in reversed(range(T_out)):
        dx = back_step(ncell - 1, t, dstack[:, :, t] + dfeed)
        for l in reversed(range(L, ncell - 1)):
            dx = back_step(l, t, dx)
        dfeed = dx  # grad wrt the decoder input of step t (= dec-last h'_{t-1}, or the encoder vector at t=0)
    denc_vec = dfeed  # grad wrt the encoder's final h (decoder input at step 0, :185)
    # (L == 1: the single decoder cell is both "first" and "last", so dfeed lands on itself.)

    # encoder, reverse time.  enc-last h'_{T_in-1} feeds the decoder input at step 0 only.
    for t in reversed(range(T_in)):
        extra_top = denc_vec if t == T_in - 1 else zeros()
        dx = back_step(L - 1, t, extra_top)
        for l in reversed(range(0, L - 1)):
            dx = back_step(l, t, dx)
    return grads


def mse_loss_and_grad(y: Tensor, target_btchw: Tensor) -> Tuple[Tensor, Tensor]:
    """training_step :55-63: y_hat = permute(y,(0,2,1,3,4)); MSELoss(mean).  Returns loss, dL/dy."""
    y_hat = y.permute(0, 2, 1, 3, 4)
    diff = y_hat - target_btchw
    loss = (diff * diff).mean()
    dy = (2.0 / diff.numel()) * diff.permute(0, 2, 1, 3, 4)
    return loss, dy.contiguous()


def rel_l2(a: Tensor, b: Tensor) -> float:
    """||a-b|| / ||b|| (b = oracle)."""
    a = a.detach().double().flatten()
    b = b.detach().double().flatten()
    den = b.norm().item()
    return (a - b).norm().item() / (den if den > 0 else 1.0)
