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


import itertools

_USE_TICK = itertools.count(1)  # orders forwards across plans (which pending graph is the oldest)


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
        # forward/backward pairing (autograd): every forward gets a new generation; a backward must present the
        # generation of the forward whose saved states are still in the workspace
        self.generation = 0
        self.pending_backward = False
        self.last_use = 0
        # gradient range statistics of the last backwards (include/clstm.h clstm_plan_grad_status), fetched without
        # synchronising: a small ring of pinned host slots, each guarded by an event
        self.overflow_policy = "raise"  # "raise" | "ignore"
        self._status_dev = None
        self._status_host = None
        self._status_slots = []  # (slot index, event, generation)
        self.last_grad_status = None

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

    def set_weights(self, params: Sequence[torch.Tensor]) -> None:
        """params in the order of include/clstm.h clstm_plan_set_weights.  Repacked on EVERY call (a few microseconds
        of device time): writes through ``p.data`` — the reference's ``init_weights`` (gan/common.py:44-50), WGAN
        weight clipping, ``load_state_dict`` — do not bump ``p._version``, so no cache key can be trusted."""
        if len(params) != self.n_params:
            raise ValueError(f"expected {self.n_params} parameter tensors, got {len(params)}")
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

    # ---- gradient range statistics ------------------------------------------------------------------------
    _STATUS_RING = 4

    def _capture_grad_status(self) -> None:
        if self._status_dev is None:
            self._status_dev = torch.empty(self._STATUS_RING, 4, dtype=torch.float32, device=self.device)
            self._status_host = torch.empty(self._STATUS_RING, 4, dtype=torch.float32).pin_memory()
        if len(self._status_slots) >= self._STATUS_RING:  # ring full: the oldest entry must be consumed first
            self.poll_grad_status(block=True)
        used = {s for s, _, _ in self._status_slots}
        slot = next(i for i in range(self._STATUS_RING) if i not in used)
        _lib.check(_lib.lib().clstm_plan_grad_status(self._h, _lib.ptr(self._status_dev[slot]), _stream_ptr(self.device)))
        self._status_host[slot].copy_(self._status_dev[slot], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._status_slots.append((slot, ev, self.generation))

    def poll_grad_status(self, block: bool = False):
        """Consume the statistics of finished backwards (``block``: wait for all of them).  Raises
        ``FloatingPointError`` when a 16-bit gradient operand overflowed (policy "raise"); returns the newest dict
        ``{"scale", "amax_dlogit", "amax_dz_scaled", "headroom_log2", "generation"}`` or None."""
        import math

        while self._status_slots:
            slot, ev, gen = self._status_slots[0]
            if block:
                ev.synchronize()
            elif not ev.query():
                break
            self._status_slots.pop(0)
            S, _, a_dl, a_dz = (float(v) for v in self._status_host[slot])
            limit = 65504.0 if self.cfg.dtype == _lib.CLSTM_F16 else 3.0e38
            st = {"scale": S, "amax_dlogit": a_dl, "amax_dz_scaled": a_dz, "generation": gen,
                  "headroom_log2": (math.log2(limit / a_dz) if 0.0 < a_dz < float("inf") else
                                    (float("-inf") if a_dz > 0.0 else float("inf")))}
            self.last_grad_status = st
            if not (a_dz < float("inf")) or not (a_dl < float("inf")):
                if self.overflow_policy == "raise":
                    what = "the loss gradient dy is not finite" if not (a_dl < float("inf")) else (
                        "a 16-bit gradient operand overflowed during back-propagation through time "
                        f"(loss scale {S:g}, max |dlogit| {a_dl:g})")
                    raise FloatingPointError(
                        f"satflow_b200: {what} in the backward of forward #{gen}; the gradients of that step are not "
                        "usable (skip the optimizer step; clip the weights / lower the learning rate, or pass a smaller "
                        "fixed grad_scale)")
        return self.last_grad_status

    def forward(self, x: torch.Tensor, y: Optional[torch.Tensor] = None, channels_last: bool = False) -> torch.Tensor:
        c = self.cfg
        if self._status_slots:
            self.poll_grad_status(block=False)  # surfaces an overflow of an earlier backward; never waits
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
        self.generation += 1
        self.last_use = next(_USE_TICK)
        self.pending_backward = self.training
        return y

    def backward(self, dy: torch.Tensor, y: torch.Tensor, grads: Sequence[Optional[torch.Tensor]], accumulate: bool = False,
                 generation: Optional[int] = None):
        """``generation``: the value of ``self.generation`` right after the forward this backward belongs to; a
        mismatch means a later forward on the same plan has overwritten the saved gates / states."""
        if generation is not None and generation != self.generation:
            raise RuntimeError(
                "satflow_b200: backward of a rollout whose saved activations were overwritten by a later forward on the "
                f"same plan (forward generation {generation}, plan is at {self.generation}).  Run backward before the "
                "next forward of the same shape, or wrap the intermediate forward in torch.no_grad()."
            )
        _require_cuda(dy, "dy")
        dy = dy.contiguous()
        y = y.contiguous()
        with torch.cuda.device(self.device):
            _lib.check(
                _lib.lib().clstm_rollout_backward(
                    self._h, _lib.ptr(dy), _lib.ptr(y), _lib.ptr_array(list(grads)), len(grads), int(accumulate),
                    _stream_ptr(self.device),
                )
            )
            if self.overflow_policy != "ignore":
                self._capture_grad_status()
        self.pending_backward = False

    INFO = {"persistent_chain": 0, "graph_enabled": 1, "graph_captures": 2, "graph_replays": 3}

    def info(self, what: str) -> int:
        """include/clstm.h clstm_plan_info: how this plan executes (persistent forward chain, CUDA-graph replay)."""
        out = ctypes.c_longlong(0)
        _lib.check(_lib.lib().clstm_plan_info(self._h, self.INFO[what], ctypes.byref(out)))
        return int(out.value)

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
    """clstm_cell_plan_t: one ConvLSTMCell.forward shape (layers/ConvLSTM.py:42-57).

    The plan's memory is two regions (include/clstm.h clstm_cell_plan_bind_split): ``scratch`` is owned here and
    reused by every call; the ``saved`` region — packed inputs, states and gates that the backward reads — belongs to
    ONE forward call.  A forward that may be differentiated gets a fresh saved region which travels with its autograd
    node, so unrolling the cell over T steps (the reference's usage, conv_lstm.py:176-196) and back-propagating
    through all of them is correct; forwards under ``torch.no_grad()`` share one."""

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
            self.saved_bytes = int(L.clstm_cell_plan_saved_bytes(self._h))
            self.scratch_bytes = int(L.clstm_cell_plan_scratch_bytes(self._h))
            self.workspace_bytes = self.saved_bytes + self.scratch_bytes
            self.scratch = _aligned_workspace(self.scratch_bytes, self.device)
        self._shared_saved = None  # saved region of the no-grad forwards
        self._bound = None  # the saved region the plan's tensor maps currently point at

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            _lib.lib().clstm_cell_plan_destroy(self._h)
            self._h = None
            self.scratch = None
            self._shared_saved = None
            self._bound = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _bind(self, saved: torch.Tensor) -> None:
        if self._bound is saved:
            return
        _lib.check(
            _lib.lib().clstm_cell_plan_bind_split(
                self._h, _lib.ptr(saved), self.saved_bytes, _lib.ptr(self.scratch), self.scratch_bytes,
                _stream_ptr(self.device),
            )
        )
        self._bound = saved

    def forward(self, x, h, c, weight, bias, keep: bool = True):
        """Returns (h_next, c_next, saved).  ``keep``: the call may be differentiated -> it gets its own saved region
        (hand it back to :meth:`backward`)."""
        for name, t in (("x", x), ("h", h), ("c", c), ("weight", weight)):
            _require_cuda(t, name)
        B, H, W, _, hid = self.shape
        # contiguous copies stay referenced until the launch is enqueued: a temporary's block could otherwise be handed
        # to the next .contiguous() by the caching allocator and two operands would alias
        xc, hc, cc, wc = x.contiguous(), h.contiguous(), c.contiguous(), weight.contiguous()
        bc = None if bias is None else bias.contiguous()
        hn = torch.empty(B, hid, H, W, dtype=torch.float32, device=x.device)
        cn = torch.empty_like(hn)
        with torch.cuda.device(self.device):
            if keep:
                saved = _aligned_workspace(self.saved_bytes, self.device)
            else:
                if self._shared_saved is None:
                    self._shared_saved = _aligned_workspace(self.saved_bytes, self.device)
                saved = self._shared_saved
            self._bind(saved)
            _lib.check(
                _lib.lib().clstm_cell_forward(
                    self._h, _lib.ptr(xc), _lib.ptr(hc), _lib.ptr(cc), _lib.ptr(wc), _lib.ptr(bc), _lib.ptr(hn),
                    _lib.ptr(cn), _stream_ptr(self.device),
                )
            )
        return hn, cn, saved

    # ---- native-layout stepping (include/clstm.h clstm_cell_native_*): inference, state kept in the plan ----
    def native_bind(self) -> None:
        if self._shared_saved is None:
            self._shared_saved = _aligned_workspace(self.saved_bytes, self.device)
        with torch.cuda.device(self.device):
            self._bind(self._shared_saved)

    def native_load(self, x=None, h=None, c=None, weight=None, bias=None, reset: bool = False) -> None:
        keep = [None if t is None else t.contiguous() for t in (x, h, c, weight, bias)]
        for name, t in zip(("x", "h", "c", "weight", "bias"), keep):
            if t is not None:
                _require_cuda(t, name)
        with torch.cuda.device(self.device):
            self.native_bind()
            _lib.check(_lib.lib().clstm_cell_native_load(self._h, *[_lib.ptr(t) for t in keep], int(reset),
                                                         _stream_ptr(self.device)))

    def native_step(self) -> None:
        with torch.cuda.device(self.device):
            self.native_bind()
            _lib.check(_lib.lib().clstm_cell_native_step(self._h, _stream_ptr(self.device)))

    def native_read(self):
        B, H, W, _, hid = self.shape
        h = torch.empty(B, hid, H, W, dtype=torch.float32, device=self.device)
        c = torch.empty_like(h)
        with torch.cuda.device(self.device):
            self.native_bind()
            _lib.check(_lib.lib().clstm_cell_native_read(self._h, _lib.ptr(h), _lib.ptr(c), _stream_ptr(self.device)))
        return h, c

    def backward(self, saved, dh, dc, weight, need_bias=True):
        B, H, W, cin, hid = self.shape
        dev = self.device
        dx = torch.empty(B, cin, H, W, dtype=torch.float32, device=dev)
        dhp = torch.empty(B, hid, H, W, dtype=torch.float32, device=dev)
        dcp = torch.empty_like(dhp)
        dw = torch.empty(weight.shape, dtype=torch.float32, device=dev)
        db = torch.empty(4 * hid, dtype=torch.float32, device=dev) if need_bias else None
        dhc = None if dh is None else dh.contiguous()
        dcc = None if dc is None else dc.contiguous()
        wc = weight.contiguous()
        with torch.cuda.device(self.device):
            self._bind(saved)
            _lib.check(
                _lib.lib().clstm_cell_backward(
                    self._h, _lib.ptr(dhc), _lib.ptr(dcc), _lib.ptr(wc), _lib.ptr(dx), _lib.ptr(dhp),
                    _lib.ptr(dcp), _lib.ptr(dw), _lib.ptr(db), _stream_ptr(self.device),
                )
            )
        return dx, dhp, dcp, dw, db
