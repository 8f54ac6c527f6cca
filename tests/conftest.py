import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_addoption(parser):
    parser.addoption("--mock-engine", action="store_true", default=False,
                     help="run the -m gpu test files on the CPU against tests/mock_engine_py.py (checks the TESTS, not the engine)")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    if config.getoption("--mock-engine"):
        import mock_engine_py
        mock_engine_py.install()


def pytest_collection_modifyitems(config, items):
    if not config.getoption("--mock-engine"):
        return
    skip = pytest.mark.skip(reason="needs the real CUDA engine (binary / multi-process / matcher library)")
    for item in items:
        if any(k in item.nodeid for k in ("test_z_adapter_gpu", "test_y_multi_gpu", "test_xx_matcher_gpu")):
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_py
    oracle_py.build()
    return oracle_py
