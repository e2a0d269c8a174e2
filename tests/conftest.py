import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run under gpurun)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")
    config.addinivalue_line("markers", "slow: oracle comparisons that need tens of seconds of CPU")


def pytest_collection_modifyitems(config, items):
    from oracle.reference_loader import reference_available

    if not reference_available():
        skip = pytest.mark.skip(reason="/root/reference not present on this box")
        for item in items:
            if "reference" in item.keywords:
                item.add_marker(skip)
    # gpu-marked tests need a B200 that libclstm accepts: skip (not fail) on a box without one
    gpu_items = [item for item in items if "gpu" in item.keywords]
    if gpu_items:
        reason = None
        try:
            from satflow_b200 import _lib

            if _lib.lib().clstm_device_check(0) != 0:
                reason = "no sm_100 device: " + _lib.lib().clstm_last_error().decode("utf-8", "replace")
        except Exception as e:  # library not built
            reason = f"libclstm.so unavailable: {e}"
        if reason:
            skip_gpu = pytest.mark.skip(reason=reason)
            for item in gpu_items:
                item.add_marker(skip_gpu)
