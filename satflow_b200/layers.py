"""ConvLSTMCell with the reference's constructor, attributes, state_dict and forward contract
(satflow/models/layers/ConvLSTM.py:7-64), executed by the fused sm_100a cell kernel through the
C ABI (clstm_cell_forward / clstm_cell_backward).  The ``conv`` sub-module is a real
``torch.nn.Conv2d`` used purely as the parameter holder, so ``conv.weight`` (4*hid, Cin+hid, kh, kw)
rows [i|f|o|g], columns [x|h] and ``conv.bias`` keep the reference layout and default init
(satflow/models/utils.py:8-10 ``get_conv_layer("standard")``).
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn as nn

from .plan import CellPlan


def get_conv_layer(conv_type: str = "standard"):
    """satflow/models/utils.py:8-20: only the "standard" branch is live on the ConvLSTM path
    ("coord" crashes in the reference's init_hidden, "3d" breaks the cat)."""
    if conv_type in ("standard", "antialiased"):  # utils.py:9-10 and :13-14 both return nn.Conv2d
        return nn.Conv2d
    if conv_type in ("coord", "3d"):
        raise ValueError(f"conv_type {conv_type!r} is not usable with ConvLSTMCell (broken in the reference as well)")
    raise ValueError(f"{conv_type} is not a recognized Conv method")  # utils.py:19


class _CellFn(torch.autograd.Function):
    """One cell step.  The activations its backward needs (packed x / h / c, gates, packed weights) live in a
    ``saved`` device region owned by THIS node, so a chain of steps through the same cell back-propagates correctly."""

    @staticmethod
    def forward(ctx, plan, x, h, c, weight, bias):
        needs_grad = any(ctx.needs_input_grad[1:])
        hn, cn, saved = plan.forward(x, h, c, weight, bias, keep=needs_grad)
        ctx.plan = plan
        ctx.has_bias = bias is not None
        ctx.saved_region = saved if needs_grad else None
        ctx.save_for_backward(weight)
        return hn, cn

    @staticmethod
    def backward(ctx, dh, dc):
        (weight,) = ctx.saved_tensors
        if ctx.saved_region is None:  # pragma: no cover - autograd only calls backward when something needed a gradient
            raise RuntimeError("satflow_b200.ConvLSTMCell: backward of a forward that ran without gradient tracking")
        dx, dhp, dcp, dw, db = ctx.plan.backward(ctx.saved_region, dh, dc, weight, need_bias=ctx.has_bias)
        ctx.saved_region = None  # single use: release the activations
        return None, dx, dhp, dcp, dw, db


class NativeCellStepper:
    """Inference stepping of one ConvLSTMCell with the recurrent state kept in the kernels' native layout (NHWC,
    16-bit h, fp32 c) inside the device plan: weights are packed once, and one ``step`` is exactly one fused kernel
    instead of pack x/h/c -> kernel -> unpack h/c (include/clstm.h "native-layout stepping").  Bit-identical to chaining
    ``ConvLSTMCell.forward`` (layers/ConvLSTM.py:42-57) under ``torch.no_grad()``; no autograd on this path.

        stepper = cell.native(batch, (H, W))          # zero state (init_hidden)
        for t in range(T):
            stepper.step(x[:, t])
        h, c = stepper.state()                        # (B, hid, H, W) float32, reference layout
    """

    def __init__(self, cell: "ConvLSTMCell", batch_size: int, image_size, device=None):
        H, W = image_size
        self.cell = cell
        dev = cell.conv.weight.device if device is None else torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("satflow_b200.ConvLSTMCell.native runs on a B200 only; move the cell to cuda first")
        self.plan = CellPlan(batch_size, H, W, cell.input_dim, cell.hidden_dim, tuple(cell.kernel_size), cell.operand_dtype, dev)
        self.refresh_weights()
        self.reset()

    def refresh_weights(self) -> None:
        """Re-pack the cell's current parameters (call after an optimizer step / load_state_dict)."""
        self.plan.native_load(weight=self.cell.conv.weight.detach(),
                              bias=None if self.cell.conv.bias is None else self.cell.conv.bias.detach())

    def reset(self, h: torch.Tensor = None, c: torch.Tensor = None) -> "NativeCellStepper":
        """Zero state (``init_hidden``), optionally seeded with reference-layout h / c."""
        self.plan.native_load(h=h, c=c, reset=True)
        return self

    def set_input(self, x: torch.Tensor) -> "NativeCellStepper":
        B, H, W, cin, _ = self.plan.shape
        if tuple(x.shape) != (B, cin, H, W):
            raise RuntimeError(f"expected input {(B, cin, H, W)}, got {tuple(x.shape)}")
        self.plan.native_load(x=x.float())
        return self

    def step(self, x: torch.Tensor = None) -> "NativeCellStepper":
        """h, c <- cell(x, (h, c)).  ``x=None`` re-uses the input packed last (e.g. timing the bare cell step)."""
        if x is not None:
            self.set_input(x)
        self.plan.native_step()
        return self

    def state(self):
        return self.plan.native_read()

    def close(self) -> None:
        self.plan.close()


class ConvLSTMCell(nn.Module):
    def __init__(self, input_dim: int, hidden_dim: int, kernel_size: Tuple[int, int], bias: bool, conv_type: str = "standard"):
        super().__init__()
        self.input_dim = input_dim
        self.hidden_dim = hidden_dim
        self.kernel_size = kernel_size
        self.padding = kernel_size[0] // 2, kernel_size[1] // 2
        self.bias = bias
        self.conv = get_conv_layer(conv_type)(
            in_channels=self.input_dim + self.hidden_dim,
            out_channels=4 * self.hidden_dim,
            kernel_size=self.kernel_size,
            padding=self.padding,
            bias=self.bias,
        )
        self.operand_dtype = "fp16"
        self._plans: Dict[tuple, CellPlan] = {}

    def _plan(self, x: torch.Tensor) -> CellPlan:
        B, C, H, W = x.shape
        key = (B, C, H, W, self.operand_dtype, x.device.index)
        plan = self._plans.get(key)
        if plan is None:
            if len(self._plans) >= 2:
                # bounded: each plan pins a scratch workspace.  The evicted plan is only dropped, not closed: an autograd
                # node of an earlier forward may still hold it (it is destroyed when the last reference goes away)
                self._plans.pop(next(iter(self._plans)))
            plan = CellPlan(B, H, W, C, self.hidden_dim, tuple(self.kernel_size), self.operand_dtype, x.device)
            self._plans[key] = plan
        return plan

    def forward(self, input_tensor: torch.Tensor, cur_state):
        h_cur, c_cur = cur_state
        if input_tensor.dim() != 4 or input_tensor.shape[1] != self.input_dim:
            raise RuntimeError(
                f"ConvLSTMCell expects input (B, {self.input_dim}, H, W), got {tuple(input_tensor.shape)}"
            )
        if not input_tensor.is_cuda:
            raise RuntimeError(
                f"satflow_b200.ConvLSTMCell runs on a B200 only: input is on {input_tensor.device}; there is no CPU path"
            )
        plan = self._plan(input_tensor)
        h_next, c_next = _CellFn.apply(plan, input_tensor, h_cur, c_cur, self.conv.weight, self.conv.bias)
        return h_next, c_next

    def native(self, batch_size: int, image_size) -> NativeCellStepper:
        """Inference stepper that keeps h / c in the device layout between steps (no per-call layout conversion)."""
        return NativeCellStepper(self, batch_size, image_size)

    def init_hidden(self, batch_size, image_size):
        height, width = image_size
        dev = self.conv.weight.device
        return (
            torch.zeros(batch_size, self.hidden_dim, height, width, device=dev),
            torch.zeros(batch_size, self.hidden_dim, height, width, device=dev),
        )
