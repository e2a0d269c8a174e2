"""Python handles over the opaque plans of libclstm.so.

A plan owns no device memory: the workspace is a torch uint8 tensor held here, so PyTorch's
caching allocator stays the single owner of HBM (include/clstm.h "Conventions").
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import torch

from . import _lib


def _stream_ptr(device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"satflow_b200: {what} must be a CUDA tensor (got device {t.device}); this package has no CPU path"
        )
    if t.dtype != torch.float32:
        raise RuntimeError(f"satflow_b200: {what} must be float32 (got {t.dtype})")


def _aligned_workspace(nbytes: int, device) -> torch.Tensor:
    buf = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
    off = (-buf.data_ptr()) % 1024
    return buf[off : off + nbytes]


class RolloutPlan:
    """clstm_plan_t: one ConvLSTM.forward shape (conv_lstm.py:205-228) on one device."""

    def __init__(
        self,
        batch: int,
        height: int,
        width: int,
        in_channels: int,
        hidden: int,
        out_channels: int,
        t_in: int,
        t_out: int,
        n_layers: int = 2,
        kernel_size=(3, 3),
        dtype: str = "fp16",
        training: bool = False,
        grad_scale: float = 0.0,
        device=None,
    ):
        L = _lib.lib()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.cfg = _lib.Config(
            batch, height, width, in_channels, hidden, out_channels, n_layers, kernel_size[0], kernel_size[1],
            t_in, t_out, _lib.DTYPES[dtype], int(training), float(grad_scale),
        )
        self.n_params = 4 * n_layers + 2
        self.training = bool(training)
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(L.clstm_plan_create(ctypes.byref(self.cfg), ctypes.byref(self._h)))
            self.workspace_bytes = int(L.clstm_plan_workspace_bytes(self._h))
            self.workspace = _aligned_workspace(self.workspace_bytes, self.device)
            _lib.check(
                L.clstm_plan_bind(self._h, _lib.ptr(self.workspace), self.workspace_bytes, _stream_ptr(self.device))
            )
        self._weights_key = None

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            _lib.lib().clstm_plan_destroy(self._h)
            self._h = None
            self.workspace = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def set_weights(self, params: Sequence[torch.Tensor], force: bool = False) -> None:
        """params in the order of include/clstm.h clstm_plan_set_weights; repacks only when they changed."""
        if len(params) != self.n_params:
            raise ValueError(f"expected {self.n_params} parameter tensors, got {len(params)}")
        key = tuple((p.data_ptr(), p._version) for p in params)
        if not force and key == self._weights_key:
            return
        for i, p in enumerate(params):
            _require_cuda(p, f"parameter {i}")
            if not p.is_contiguous():
                raise RuntimeError("satflow_b200: parameters must be contiguous")
        with torch.cuda.device(self.device):
            _lib.check(
                _lib.lib().clstm_plan_set_weights(
                    self._h, _lib.ptr_array(list(params)), len(params), _stream_ptr(self.device)
                )
            )
        self._weights_key = key

    def forward(self, x: torch.Tensor, y: Optional[torch.Tensor] = None, channels_last: bool = False) -> torch.Tensor:
        c = self.cfg
        _require_cuda(x, "x")
        want = ((c.batch, c.t_in, c.height, c.width, c.in_channels) if channels_last
                else (c.batch, c.t_in, c.in_channels, c.height, c.width))
        if tuple(x.shape) != want:
            raise ValueError(f"x has shape {tuple(x.shape)}, plan expects {want}")
        x = x.contiguous()
        if y is None:
            y = torch.empty(c.batch, c.out_channels, c.t_out, c.height, c.width, dtype=torch.float32, device=x.device)
        with torch.cuda.device(self.device):
            _lib.check(
                _lib.lib().clstm_rollout_forward_layout(
                    self._h, _lib.ptr(x), 1 if channels_last else 0, _lib.ptr(y), _stream_ptr(self.device)
                )
            )
        return y

    def backward(self, dy: torch.Tensor, y: torch.Tensor, grads: Sequence[Optional[torch.Tensor]], accumulate: bool = False):
        _require_cuda(dy, "dy")
        dy = dy.contiguous()
        with torch.cuda.device(self.device):
            _lib.check(
                _lib.lib().clstm_rollout_backward(
                    self._h, _lib.ptr(dy), _lib.ptr(y), _lib.ptr_array(list(grads)), len(grads), int(accumulate),
                    _stream_ptr(self.device),
                )
            )

    KERNELS = {"cell_fwd": 0, "gate_grad": 1, "dgrad": 2, "wgrad": 3, "wgrad+gate_grad": 4, "dgrad_fused": 5}

    def profile_kernel(self, kind: str, cell: int, step: int) -> None:
        """Re-launch one kernel of (cell, step) (measurement hook, include/clstm.h clstm_plan_profile_kernel)."""
        with torch.cuda.device(self.device):
            _lib.check(
                _lib.lib().clstm_plan_profile_kernel(self._h, self.KERNELS[kind], cell, step, _stream_ptr(self.device))
            )

    def profile_cell_step(self, cell: int, step: int) -> None:
        self.profile_kernel("cell_fwd", cell, step)

    def read_state(self, cell: int, step: int):
        c = self.cfg
        h = torch.empty(c.batch, c.hidden, c.height, c.width, dtype=torch.float32, device=self.device)
        cc = torch.empty_like(h)
        with torch.cuda.device(self.device):
            _lib.check(
                _lib.lib().clstm_plan_read_state(self._h, cell, step, _lib.ptr(h), _lib.ptr(cc), _stream_ptr(self.device))
            )
        return h, cc


class CellPlan:
    """clstm_cell_plan_t: one ConvLSTMCell.forward shape (layers/ConvLSTM.py:42-57)."""

    def __init__(self, batch, height, width, in_channels, hidden, kernel_size=(3, 3), dtype="fp16", device=None):
        L = _lib.lib()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.shape = (batch, height, width, in_channels, hidden)
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(
                L.clstm_cell_plan_create(
                    batch, height, width, in_channels, hidden, kernel_size[0], kernel_size[1], _lib.DTYPES[dtype],
                    ctypes.byref(self._h),
                )
            )
            self.workspace_bytes = int(L.clstm_cell_plan_workspace_bytes(self._h))
            self.workspace = _aligned_workspace(self.workspace_bytes, self.device)
            _lib.check(
                L.clstm_cell_plan_bind(self._h, _lib.ptr(self.workspace), self.workspace_bytes, _stream_ptr(self.device))
            )

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            _lib.lib().clstm_cell_plan_destroy(self._h)
            self._h = None
            self.workspace = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def forward(self, x, h, c, weight, bias):
        for name, t in (("x", x), ("h", h), ("c", c), ("weight", weight)):
            _require_cuda(t, name)
        B, H, W, _, hid = self.shape
        hn = torch.empty(B, hid, H, W, dtype=torch.float32, device=x.device)
        cn = torch.empty_like(hn)
        with torch.cuda.device(self.device):
            _lib.check(
                _lib.lib().clstm_cell_forward(
                    self._h, _lib.ptr(x.contiguous()), _lib.ptr(h.contiguous()), _lib.ptr(c.contiguous()),
                    _lib.ptr(weight.contiguous()), _lib.ptr(None if bias is None else bias.contiguous()),
                    _lib.ptr(hn), _lib.ptr(cn), _stream_ptr(self.device),
                )
            )
        return hn, cn

    def backward(self, dh, dc, weight, need_bias=True):
        B, H, W, cin, hid = self.shape
        dev = self.device
        dx = torch.empty(B, cin, H, W, dtype=torch.float32, device=dev)
        dhp = torch.empty(B, hid, H, W, dtype=torch.float32, device=dev)
        dcp = torch.empty_like(dhp)
        dw = torch.empty_like(weight)
        db = torch.empty(4 * hid, dtype=torch.float32, device=dev) if need_bias else None
        with torch.cuda.device(self.device):
            _lib.check(
                _lib.lib().clstm_cell_backward(
                    self._h, _lib.ptr(None if dh is None else dh.contiguous()),
                    _lib.ptr(None if dc is None else dc.contiguous()), _lib.ptr(weight), _lib.ptr(dx), _lib.ptr(dhp),
                    _lib.ptr(dcp), _lib.ptr(dw), _lib.ptr(db), _stream_ptr(self.device),
                )
            )
        return dx, dhp, dcp, dw, db
