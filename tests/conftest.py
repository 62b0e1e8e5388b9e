import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def walk_cfg():
    from samcon_b200.config import Config
    return Config("walk")


@pytest.fixture(scope="session")
def character(walk_cfg):
    from samcon_b200.model import Character
    return Character.from_config(walk_cfg)


@pytest.fixture(scope="session")
def oracle(character):
    from oracle.oracle import Oracle
    return Oracle(character)


@pytest.fixture(scope="session")
def pool(character):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from samcon_b200.pool import GpuSamplePool
    p = GpuSamplePool(character, 0)
    yield p
    p.close()
