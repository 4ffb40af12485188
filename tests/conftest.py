import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "t-route_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    for item in items:
        if item.get_closest_marker("gpu") is not None and item.get_closest_marker("timeout") is None:
            # a hung kernel must end the run, not the GPU box's lease (thread method: the process is killed even when the
            # main thread sits inside cudaDeviceSynchronize)
            item.add_marker(pytest.mark.timeout(1200, method="thread"))


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure), built on demand."""
    from oracle import oracle as o
    o.build()
    o.lib()
    return o
