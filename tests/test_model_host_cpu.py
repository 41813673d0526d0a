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


def test_overlap_schedule_is_taken_and_changes_nothing():
    """settings.overlap_root_update: with the side-stream helpers replaced by inline stand-ins
    (cpu_ops_mock) the pre-started inverse-root update and the side-section projection run on the CPU; the stream of
    (rmse, nll, loss) must equal the serial schedule's exactly and the pre-start must really have been used."""
    import warnings

    import torch

    import online_gp_b200.settings as S
    from online_gp_b200.lazy.updated_root_lazy_tensor import UpdatedRootLazyTensor
    from online_gp_b200.models import OnlineSKIRegression
    from online_gp_b200.models.stems import Identity

    d, g, n0, steps = 2, 24, 40, 4
    gen = torch.Generator().manual_seed(17)
    X = torch.rand(n0 + steps, d, generator=gen) * 2 - 1
    y = (torch.sin(3 * X.sum(-1)) + 0.1 * torch.randn(n0 + steps, generator=gen)).unsqueeze(-1)
    started = []
    orig = UpdatedRootLazyTensor.prestart_update_sparse

    def spy(self, idx, vval):
        ok = orig(self, idx, vval)
        started.append(ok)
        return ok

    UpdatedRootLazyTensor.prestart_update_sparse = spy
    try:
        runs = {}
        for on in (True, False):
            with warnings.catch_warnings(), S.max_cholesky_size(0), S.max_root_decomposition_size(32), \
                    S.overlap_root_update(on):
                warnings.simplefilter("ignore")
                reg = OnlineSKIRegression(Identity(d), X[:n0], y[:n0], lr=1e-2, grid_size=g, grid_bound=1.0)
                rows = []
                for t in range(steps):
                    xt, yt = X[n0 + t:n0 + t + 1], y[n0 + t:n0 + t + 1]
                    rmse, nll = reg.evaluate(xt, yt)
                    _, loss = reg.update(xt, yt)
                    rows.append((rmse, nll, loss))
                wtw = reg.gp._kernel_cache["WtW"]
                assert getattr(wtw, "_pending", None) is None
                runs[on] = (rows, wtw.root.clone(), wtw.inv_root.clone())
            if on:
                assert started == [True] * steps, started
        assert len(started) == steps                                   # the serial run never pre-starts
        assert runs[True][0] == runs[False][0]
        assert torch.equal(runs[True][1], runs[False][1]) and torch.equal(runs[True][2], runs[False][2])
    finally:
        UpdatedRootLazyTensor.prestart_update_sparse = orig
