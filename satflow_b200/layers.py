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
            if len(self._plans) >= 2:  # bounded: each plan pins a workspace
                self._plans.pop(next(iter(self._plans))).close()
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

    def init_hidden(self, batch_size, image_size):
        height, width = image_size
        dev = self.conv.weight.device
        return (
            torch.zeros(batch_size, self.hidden_dim, height, width, device=dev),
            torch.zeros(batch_size, self.hidden_dim, height, width, device=dev),
        )
