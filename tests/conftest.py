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
    config.addinivalue_line("markers", "first_hw_run: exercises code that has not yet executed on a GPU; collected last so "
                                       "that under `-x` it cannot mask the tests of already verified code")


def pytest_collection_modifyitems(config, items):
    late = [it for it in items if it.get_closest_marker("first_hw_run")]
    if late:
        items[:] = [it for it in items if not it.get_closest_marker("first_hw_run")] + late


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
