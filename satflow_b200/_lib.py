"""ctypes binding of libclstm.so (the C ABI in include/clstm.h).

There is deliberately no fallback: if the shared library is missing or no sm_100 device is
present, every product entry point raises.  ``oracle/`` is never imported from here.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_size_t, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CLSTM_LIB", os.path.join(_HERE, "libclstm.so"))  # CLSTM_LIB: A/B experiments only

CLSTM_F16 = 0
CLSTM_BF16 = 1
DTYPES = {"fp16": CLSTM_F16, "f16": CLSTM_F16, "float16": CLSTM_F16, "bf16": CLSTM_BF16, "bfloat16": CLSTM_BF16}


class ClstmError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libclstm error {code}: {message}")
        self.code = code


class Config(ctypes.Structure):
    """Mirror of clstm_config_t (include/clstm.h)."""

    _fields_ = [
        ("batch", c_int32),
        ("height", c_int32),
        ("width", c_int32),
        ("in_channels", c_int32),
        ("hidden", c_int32),
        ("out_channels", c_int32),
        ("n_layers", c_int32),
        ("kernel_h", c_int32),
        ("kernel_w", c_int32),
        ("t_in", c_int32),
        ("t_out", c_int32),
        ("dtype", c_int32),
        ("training", c_int32),
        ("grad_scale", c_float),
    ]


_PROTOTYPES = {
    "clstm_last_error": (c_char_p, []),
    "clstm_abi_version": (c_int, []),
    "clstm_device_check": (c_int, [c_int]),
    "clstm_plan_create": (c_int, [POINTER(Config), POINTER(c_void_p)]),
    "clstm_plan_destroy": (c_int, [c_void_p]),
    "clstm_plan_workspace_bytes": (c_size_t, [c_void_p]),
    "clstm_plan_bind": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "clstm_plan_set_weights": (c_int, [c_void_p, POINTER(c_void_p), c_int, c_void_p]),
    "clstm_rollout_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "clstm_rollout_forward_layout": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "clstm_rollout_backward": (c_int, [c_void_p, c_void_p, c_void_p, POINTER(c_void_p), c_int, c_int, c_void_p]),
    "clstm_plan_grad_status": (c_int, [c_void_p, c_void_p, c_void_p]),
    "clstm_plan_info": (c_int, [c_void_p, c_int, POINTER(ctypes.c_longlong)]),
    "clstm_plan_read_state": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "clstm_plan_profile_kernel": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p]),
    "clstm_cell_plan_create": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, POINTER(c_void_p)]),
    "clstm_cell_plan_destroy": (c_int, [c_void_p]),
    "clstm_cell_plan_workspace_bytes": (c_size_t, [c_void_p]),
    "clstm_cell_plan_bind": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "clstm_cell_plan_saved_bytes": (c_size_t, [c_void_p]),
    "clstm_cell_plan_scratch_bytes": (c_size_t, [c_void_p]),
    "clstm_cell_plan_bind_split": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_void_p]),
    "clstm_cell_forward": (c_int, [c_void_p] + [c_void_p] * 7 + [c_void_p]),
    "clstm_cell_backward": (c_int, [c_void_p] + [c_void_p] * 8 + [c_void_p]),
    "clstm_cell_native_load": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "clstm_cell_native_step": (c_int, [c_void_p, c_void_p]),
    "clstm_cell_native_read": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "clstm_mse_loss_grad": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "clstm_launch_count": (c_uint64, []),
    "clstm_trace_enable": (c_int, [c_int]),
    "clstm_trace_report": (ctypes.c_longlong, [ctypes.c_char_p, c_size_t]),
}

_lib = None


def lib() -> ctypes.CDLL:
    """Loads libclstm.so (built in-tree by ``__graft_entry__.build()``); raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise ClstmError(
                -100,
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  There is no CPU / eager fallback for this path.",
            )
        handle = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in _PROTOTYPES.items():
            fn = getattr(handle, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def exported_symbols():
    return sorted(_PROTOTYPES)


def check(rc: int) -> None:
    if rc != 0:
        raise ClstmError(rc, lib().clstm_last_error().decode("utf-8", "replace"))


def ptr(t) -> c_void_p:
    """Device pointer of a torch tensor (or NULL for None)."""
    return c_void_p(0 if t is None else t.data_ptr())


def ptr_array(tensors):
    arr = (c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr


def trace_enable(capacity: int = 4096) -> None:
    """Record one CUDA event after every library launch (0 disables); see clstm_trace_enable in include/clstm.h."""
    check(lib().clstm_trace_enable(int(capacity)))


def trace_report() -> str:
    """Per-kernel in-situ timing table of the launches since trace_enable / the last report."""
    buf = ctypes.create_string_buffer(1 << 16)
    n = lib().clstm_trace_report(buf, len(buf))
    if n < 0:
        check(int(n))
    return buf.value.decode("utf-8", "replace")
