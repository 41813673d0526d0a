"""GPU validation of the "next" rows of SURVEY §8f: every case of tests/next_rows_cases.py on cuda:0 through the real
kernels (batched posteriors over candidate sets, fantasies, BO / active-learning loops in miniature, fit(), trainable stem)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

import next_rows_cases
from next_rows_cases import *  # noqa: F401,F403

next_rows_cases.DEV = "cuda:0"
next_rows_cases.TOLX = 1e3


@pytest.fixture(autouse=True)
def _fp64_default_gpu():
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    yield
    torch.set_default_dtype(prev)
