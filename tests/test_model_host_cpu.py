"""CPU: host logic of the model layer (operator algebra, cache protocol, MLL assembly, autograd wiring, wrapper
loop) on the cases of tests/model_cases.py, with the CUDA entry points mocked by the oracle (tests/cpu_ops_mock.py).
The kernels themselves are checked by the -m gpu suite."""
import pytest

import cpu_ops_mock
import model_cases
from model_cases import *  # noqa: F401,F403


@pytest.fixture(autouse=True)
def _mock_ops():
    model_cases.DEV = "cpu"
    with cpu_ops_mock.install():
        yield
    model_cases.DEV = "cuda:0"
