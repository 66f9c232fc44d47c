import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: full-size parity (about a minute of CPU reference work per case)")


@pytest.fixture(scope="session")
def lib():
    from vidseg_diffusion_b200 import _lib
    return _lib.load()


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("this test is marked gpu and needs a CUDA device")
    return torch.device("cuda", 0)


@pytest.fixture(params=[0, 1], ids=["pair16", "packed8"])
def operand_mode(request, lib):
    """Runs a test under both operand policies of the tensor-core GEMMs (include/vidseg_b200.h,
    vidseg_set_operand_mode): 0 = fp16 pairs everywhere (3 MMAs per product), 1 = fp16 + fp8 corrections (default)."""
    lib.vidseg_set_operand_mode(request.param)
    yield request.param
    lib.vidseg_set_operand_mode(1)
