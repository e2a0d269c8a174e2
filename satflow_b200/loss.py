"""Fused MSE loss of EncoderDecoderConvLSTM.training_step (conv_lstm.py:55-69): loss, its gradient and the
per-frame losses in one pass over y and the target (C ABI clstm_mse_loss_grad)."""
from __future__ import annotations

import ctypes

import torch

from . import _lib


class _FusedMSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y: torch.Tensor, target: torch.Tensor):
        B, C, T, H, W = y.shape
        if tuple(target.shape) != (B, T, C, H, W):
            raise ValueError(f"target must be (B,T,C,H,W) = {(B, T, C, H, W)}, got {tuple(target.shape)}")
        y = y.contiguous()
        target = target.contiguous()
        need_grad = ctx.needs_input_grad[0]
        dy = torch.empty_like(y) if need_grad else None
        partial = torch.empty(B * T * C, dtype=torch.float32, device=y.device)
        out = torch.empty(1 + T, dtype=torch.float32, device=y.device)
        with torch.cuda.device(y.device):
            _lib.check(
                _lib.lib().clstm_mse_loss_grad(
                    _lib.ptr(y), _lib.ptr(target), B, C, T, H, W, _lib.ptr(dy), _lib.ptr(partial), _lib.ptr(out),
                    ctypes.c_void_p(torch.cuda.current_stream(y.device).cuda_stream),
                )
            )
        ctx.save_for_backward(dy)
        loss, frames = out[0], out[1:]
        ctx.mark_non_differentiable(frames)
        return loss, frames

    @staticmethod
    def backward(ctx, g_loss, g_frames):
        (dy,) = ctx.saved_tensors
        return dy * g_loss, None


def fused_mse(y: torch.Tensor, target: torch.Tensor):
    """y: (B, C, T, H, W) rollout output; target: (B, T, C, H, W).  Returns (mean loss, per-frame losses [T]) —
    numerically MSELoss()(y.permute(0, 2, 1, 3, 4), target) and its per-frame values."""
    if not (y.is_cuda and target.is_cuda and y.dtype == torch.float32 and target.dtype == torch.float32):
        raise RuntimeError("fused_mse needs float32 CUDA tensors (there is no CPU path)")
    return _FusedMSE.apply(y, target)
