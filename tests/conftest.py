import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.build()
    return o


def pytest_sessionfinish(session, exitstatus):
    """with the bounds-checked library (SNAQ_B200_LIB=.../libsnaqb200_dbg.so) report the device-side index checks of the run"""
    if not os.environ.get("SNAQ_B200_LIB"):
        return
    try:
        import snaq_jl_b200 as S
        eng = S.Engine(0)
        checks, bad = eng.debug_bounds()
        print(f"\n[snaq bounds-checked build] {checks} index checks, {bad} violations")
        if bad > 0:
            session.exitstatus = 1
    except Exception as exc:  # noqa: BLE001
        print(f"\n[snaq bounds-checked build] could not read the counters: {exc}")
