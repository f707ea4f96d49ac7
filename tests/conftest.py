import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Builds everything that is missing (cheap no-op when up to date)."""
    need = [
        os.path.join(ROOT, "zra_b200", "libzra_b200.so"),
        os.path.join(ROOT, "zra_b200", "libzra_synth.so"),
        os.path.join(ROOT, "oracle", "libzra_oracle.so"),
        os.path.join(ROOT, "tests", "host_sim", "libsim_decode.so"),
    ]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__

        __graft_entry__.build()
