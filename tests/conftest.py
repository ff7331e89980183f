import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _guard_report(c, name):
    """SYNSEG_GUARD build (tools/guard_run.sh): canary zones compared / damaged for one context; fails the run on damage."""
    import ctypes as C
    zones, guard = C.c_int64(), C.c_int32()
    bad = c.lib.synseg_guard_violations(c._h, C.byref(zones), C.byref(guard))
    if guard.value:
        print(f"\nSYNSEG_GUARD [{name}]: {zones.value} canary zones compared, {bad} damaged")
        assert bad == 0, f"{bad} scratch canaries were overwritten"


@pytest.fixture(scope="session")
def ctx():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from synapta_image_segmentation_b200.ops import Context
    c = Context(0)
    yield c
    _guard_report(c, "test context")
    from synapta_image_segmentation_b200 import detector
    for dev, other in detector._contexts.items():
        if other._h:
            _guard_report(other, f"process-wide context of cuda:{dev}")
    c.close()
