"""CPU (kernels mocked by the oracle, tests/cpu_ops_mock.py): host logic of the "next" rows of SURVEY §8f — the cases of
``tests/next_rows_cases.py`` (the same cases run on the real kernels in ``test_next_rows_gpu.py``), plus the Dirichlet
classifier wrapper (out of the hot-path scope, CPU only)."""
import warnings

import pytest
import torch

import cpu_ops_mock
import next_rows_cases
from next_rows_cases import *  # noqa: F401,F403
from next_rows_cases import _model  # noqa: F401


@pytest.fixture(autouse=True)
def _mock_ops():
    next_rows_cases.DEV = "cpu"
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    with cpu_ops_mock.install():
        yield
    torch.set_default_dtype(prev)
    next_rows_cases.DEV = "cuda:0"


def _two_blobs(n, gen):
    y = (torch.rand(n, generator=gen) > 0.5).long()
    centers = torch.tensor([[-0.8, -0.5], [0.7, 0.6]])
    x = centers[y] + 0.35 * torch.randn(n, 2, generator=gen)
    return x.clamp(-2.9, 2.9), y


def test_dirichlet_ski_classifier_batch_and_online():
    """tests/classification/test_ski_classifier.py (batch fit with an Identity stem, then streaming updates) on a
    synthetic two-class problem (the Banana data set of the reference is not available offline)."""
    from online_gp_b200 import settings as S
    from online_gp_b200.models import OnlineSKIClassifier
    from online_gp_b200.models.online_ski_classifier import dirichlet_transform
    from online_gp_b200.models.stems import Identity
    gen = torch.Generator().manual_seed(0)
    y_t, alpha, s2 = dirichlet_transform(torch.tensor([0, 1, 1]), 1e-2)
    assert y_t.shape == (3, 2) and torch.allclose(alpha, torch.tensor([[1.01, 0.01], [0.01, 1.01], [0.01, 1.01]]))
    assert torch.allclose(s2, torch.log(1.0 / alpha + 1.0)) and bool((y_t[:, 0] > y_t[:, 1]).tolist() == [True, False, False])
    train_x, train_y = _two_blobs(120, gen)
    test_x, test_y = _two_blobs(200, gen)
    with warnings.catch_warnings(), S.max_root_decomposition_size(512), S.max_cholesky_size(2048):
        warnings.simplefilter("ignore")
        clf = OnlineSKIClassifier(Identity(2), train_x[:60], train_y[:60], 1e-2, 1e-1, 12, 3.1)
        rec = clf.fit(train_x[:60], train_y[:60], 8)
        assert len(rec) == 8 and rec[-1]["epoch"] == 8
        acc0 = clf.predict(test_x).eq(test_y).float().mean().item()
        assert acc0 >= 0.9
        n_before = clf.gp.num_data
        for t in range(60, 120, 10):
            stem_loss, gp_loss = clf.update(train_x[t:t + 10], train_y[t:t + 10])
            assert stem_loss == 0 and gp_loss == gp_loss          # Identity stem: nothing to train
        assert clf.gp.num_data == n_before + 60
        acc1 = clf.predict(test_x).eq(test_y).float().mean().item()
        assert acc1 >= 0.9


def test_dirichlet_ski_classifier_online_learned_features():
    """Streaming updates with a trainable stem (LinearStem: Linear + BatchNorm + tanh): the stem step goes through the
    Sherman-Morrison partial MLL with two outputs (online_ski_classifier.py:103-117)."""
    from online_gp_b200 import settings as S
    from online_gp_b200.models import OnlineSKIClassifier
    from online_gp_b200.models.stems import LinearStem
    gen = torch.Generator().manual_seed(1)
    torch.manual_seed(1)
    train_x, train_y = _two_blobs(40, gen)
    with warnings.catch_warnings(), S.max_root_decomposition_size(512), S.max_cholesky_size(2048):
        warnings.simplefilter("ignore")
        clf = OnlineSKIClassifier(LinearStem(2, 2).double(), train_x[:8], train_y[:8], 1e-2, 1e-3, 10, 1.0)
        clf.set_lr(1e-3, 1e-3, bn_mom=0.05)
        clf.eval()                          # as after fit(): BatchNorm uses its running statistics for single points
        w0 = clf.stem[0].weight.detach().clone()
        for t in range(8, 40):              # one point per update: the Sherman-Morrison increment is a rank-one formula
            stem_loss, gp_loss = clf.update(train_x[t:t + 1], train_y[t:t + 1])
            assert stem_loss == stem_loss and gp_loss == gp_loss and abs(stem_loss) < 1e6
        assert clf.gp.num_data == 40
        assert not torch.equal(w0, clf.stem[0].weight.detach())           # the stem really trained
        assert clf.predict(train_x).shape == (40,)


