import sys
from pathlib import Path

import pytest

import os

# The device-group tests put up to 8 ranks of a sharded frame on ONE GPU.  The ranks wait for each other inside tiny kernels
# (k_shard_sync, k_display_spin); with CUDA's default of 8 hardware work queues per device two ranks' streams can share a
# queue, and a waiting kernel at its head then blocks the very kernel it waits for until the 4 s time-out.  On a real
# multi-GPU box every rank has its own device and queues.  Must be set before the CUDA context exists.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "rendering-fw_b200" / "python"))
sys.path.insert(0, str(REPO))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built():
    """Build the product library and the oracle once per session (no-op when up to date)."""
    import __graft_entry__ as g

    g.build()
    return True


@pytest.fixture(scope="session")
def oracle_lib(built):
    from oracle.oracle_lib import load_oracle

    return load_oracle()


@pytest.fixture(scope="session")
def product_lib(built):
    import rfwb200 as R

    return R.load_product()
