import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# Slab-group tests put several slabs on ONE GPU; a slab's one-warp reduction kernel spins
# on its peers' mailboxes, so the peers' streams must not share a hardware queue with it
# (read by the CUDA driver at initialisation; inherited by the subprocesses tests start).
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
# Small single-GPU solves run as ONE persistent cooperative kernel by default (k_cg_persistent).
# Almost every parity case here IS small, and the kernels that matter at 512^3 (TMA-staged
# stencil, streaming update, CUDA graph) must keep their coverage: the suite runs with the
# persistent loop off, and tests/test_gpu_zz_round2.py switches it on for its own battery.
os.environ.setdefault("APHCG_PERSISTENT", "0")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Everything compiled (CUDA library, oracle)."""
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope="session")
def gpu(built):
    from aphros_b200 import capi
    if capi.device_count() < 1:
        pytest.fail("GPU test selected but no CUDA device is visible: the product has no CPU path")
    return True
