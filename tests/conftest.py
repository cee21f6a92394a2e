import os
import sys

import pytest

import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# the dataset layer's binary cache (modules/load/fast.py) must not land in the user's ~/.cache during test runs
os.environ.setdefault("OEA_CACHE_DIR", os.path.join(tempfile.gettempdir(), "oea_dataset_cache_tests_%d" % os.getpid()))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "first_hw_run: exercises code that has not yet executed on a GPU; collected last and "
                                       "quarantined (a failure is reported as xfailed + listed in the summary)")


def pytest_collection_modifyitems(config, items):
    """Tests of code that has never executed on a GPU (`first_hw_run`) are collected last and QUARANTINED: they run, but
    a failure is reported as xfailed (and listed with its first line in the terminal summary) instead of turning the
    suite of hardware-verified tests red or — under `-x` — cutting it short; a stuck kernel ends the process after 10
    minutes instead of holding the GPU box.  OEA_STRICT_FIRST_HW=1 makes them ordinary tests.  A test leaves the
    quarantine (the marker is removed) after its first green run on hardware."""
    late = [it for it in items if it.get_closest_marker("first_hw_run")]
    if late:
        items[:] = [it for it in items if not it.get_closest_marker("first_hw_run")] + late
    if os.environ.get("OEA_STRICT_FIRST_HW") == "1":
        return
    for it in late:
        it.add_marker(pytest.mark.xfail(strict=False, reason="first run on hardware (quarantined, see tests/conftest.py)"))
        if config.pluginmanager.hasplugin("timeout"):
            it.add_marker(pytest.mark.timeout(600, method="thread"))


def pytest_terminal_summary(terminalreporter, exitstatus, config):
    xfailed = [r for r in terminalreporter.stats.get("xfailed", []) if "first run on hardware" in str(getattr(r, "wasxfail", ""))]
    xpassed = [r for r in terminalreporter.stats.get("xpassed", [])]
    if not xfailed and not xpassed:
        return
    terminalreporter.section("first hardware runs (quarantined)")
    terminalreporter.write_line("%d passed on their first hardware run, %d FAILED:" % (len(xpassed), len(xfailed)))
    for r in xfailed:
        crash = getattr(r.longrepr, "reprcrash", None)
        text = str(r.longrepr).strip().splitlines()
        where = "%s:%s: %s" % (os.path.basename(crash.path), crash.lineno, crash.message.splitlines()[0]) if crash else \
            (text[-1] if text else "")
        terminalreporter.write_line("  FAILED %s :: %s" % (r.nodeid, where[:300]))


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
