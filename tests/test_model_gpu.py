"""GPU parity of the model layer: runs every case of tests/model_cases.py on cuda:0 through the real kernels."""
import pytest

pytestmark = pytest.mark.gpu

import model_cases
from model_cases import *  # noqa: F401,F403

model_cases.DEV = "cuda:0"
