import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def golden(*parts):
    return os.path.join(GOLDEN, *parts)


@pytest.fixture(scope='session')
def oracle_lib():
    import build
    build.build_oracle()
    import oracle
    return oracle.lib()


@pytest.fixture(scope='session')
def gpu_ctx():
    """The CUDA path must be the one that runs: no skip, no fallback when the extension is missing."""
    import torch
    assert torch.cuda.is_available(), 'gpu-marked test run without a CUDA device'
    from dynast_b200._lib import Context
    return Context.get(0)


@pytest.fixture
def fake_gpu(monkeypatch, oracle_lib):
    """Host-logic tests: numpy stand-ins for the device stages (tests/fake_device.py). CPU only."""
    import fake_device
    fake_device.install(monkeypatch)
    return True


@pytest.fixture(params=['cpu-host-logic', pytest.param('gpu', marks=pytest.mark.gpu)])
def backend(request):
    """Operator tests run twice: against the real kernels (-m gpu) and, for the host logic alone, against the
    numpy stand-ins (-m 'not gpu')."""
    if request.param == 'gpu':
        request.getfixturevalue('gpu_ctx')
    else:
        request.getfixturevalue('fake_gpu')
    return request.param
