"""``OnlineSKIClassifier`` — binary classification on top of the WISKI core through the Dirichlet / log-normal target
transform (Milios et al. 2018): class labels become two regression outputs with per-point heteroskedastic noise,
which is exactly the (targets [n, t], fixed noise [n, t]) form ``FixedNoiseOnlineSKIGP`` consumes with t = 2.

Same classes, constructor arguments and methods as ``online_gp/models/gp_dirichlet_classification.py:5-45`` and
``online_gp/models/online_ski_classifier.py:13-145`` (``fit``, ``update``, ``predict``, ``set_lr``, ``set_train_data``);
nothing here adds arithmetic of its own — every panel / Kronecker / interpolation operation goes through the same
CUDA kernels as the regression path.
"""
import torch
from torch.optim.lr_scheduler import CosineAnnealingLR

from .. import settings
from ..mlls.batched_woodbury_marginal_log_likelihood import BatchedWoodburyMarginalLogLikelihood
from ..mlls.streaming_partial_mll import sm_partial_mll
from ..settings import detach_interp_coeff
from .batched_fixed_noise_online_gp import FixedNoiseOnlineSKIGP


def dirichlet_transform(labels, alpha_eps, num_classes=2):
    """labels [n] (int) -> (regression targets [n, C], alpha [n, C], noise variances [n, C]):
    alpha = alpha_eps + onehot,  sigma^2 = log(1 / alpha + 1),  y = log(alpha) - sigma^2 / 2."""
    n = labels.size(0)
    alpha = alpha_eps * torch.ones(n, num_classes, device=labels.device)
    alpha[torch.arange(n, device=labels.device), labels] += 1
    sigma2 = torch.log(1.0 / alpha + 1.0)
    return alpha.log() - 0.5 * sigma2, alpha, sigma2


class DirichletGPClassifier(torch.nn.Module):
    def __init__(self, stem, gp, mll, alpha_eps, lr, *args, **kwargs):
        super().__init__()
        self.stem, self.gp, self.mll, self.alpha_eps = stem, gp, mll, alpha_eps
        self.optimizer = torch.optim.Adam(self.parameters(), lr=lr)
        self._target_batch_shape = []

    def _transform_targets(self, targets, alpha_eps):
        return dirichlet_transform(targets, alpha_eps)

    def forward(self, inputs):
        return self.gp(self.stem(inputs.view(-1, self.stem.input_dim)))

    def predict(self, inputs):
        self.eval()
        return self(inputs).mean.argmax(0)             # outputs are the leading (batch) dimension of the MVN

    def fit(self, inputs, targets, num_epochs):
        raise NotImplementedError

    def set_train_data(self, inputs, targets, noise):
        self.gp.set_train_data(inputs, targets, noise)

    def set_lr(self, gp_lr, stem_lr=None):
        self.optimizer = torch.optim.Adam([dict(params=self.gp.parameters(), lr=gp_lr),
                                           dict(params=self.stem.parameters(), lr=gp_lr if stem_lr is None else stem_lr)])


class OnlineSKIClassifier(DirichletGPClassifier):
    def __init__(self, stem, init_x, init_y, alpha_eps, lr, grid_size, grid_bound, **kwargs):
        stem = stem.to(init_x.device)
        y0, _, noise0 = dirichlet_transform(init_y, alpha_eps)
        gp = FixedNoiseOnlineSKIGP(
            stem(init_x).detach(), y0, noise0,
            grid_bounds=torch.tensor([[-grid_bound, grid_bound]] * stem.output_dim),
            grid_size=[grid_size] * stem.output_dim,
        )
        super().__init__(stem, gp, BatchedWoodburyMarginalLogLikelihood(gp.likelihood, gp), alpha_eps, lr)
        del self.optimizer
        self._make_optimizers(lr, lr)
        self._target_batch_shape = torch.Size([y0.shape[-1]]) if y0.shape[-1] != 1 else torch.Size()
        self._raw_inputs = [init_x]

    def _make_optimizers(self, gp_lr, stem_lr):
        self.gp_optimizer = torch.optim.Adam(self.gp.parameters(), lr=gp_lr)
        self.stem_optimizer = torch.optim.Adam(self.stem.parameters(), lr=stem_lr)

    # ---- batch (pre-)training: the caches are rebuilt from the current features after every step
    def _refresh_features(self, inputs, targets):
        features = self.stem(inputs)
        y, _, noise = dirichlet_transform(targets, self.alpha_eps)
        self.set_train_data(features, y, noise)
        self.gp.zero_grad()
        return features

    def fit(self, inputs, targets, num_epochs, test_dataset=None):
        scheds = [CosineAnnealingLR(opt, num_epochs, 1e-4) for opt in (self.gp_optimizer, self.stem_optimizer)]
        features = self._refresh_features(inputs, targets)
        records = []
        for epoch in range(num_epochs):
            self.train()
            self.mll.train()
            for opt in (self.gp_optimizer, self.stem_optimizer):
                opt.zero_grad()
            loss = -self.mll(self.gp(features), targets).sum()
            loss.backward()
            for step in (self.gp_optimizer, self.stem_optimizer, *scheds):
                step.step()
            features = self._refresh_features(inputs, targets)
            acc = float("NaN")
            if test_dataset is not None:
                test_x, test_y = test_dataset[:]
                acc = self.predict(test_x).eq(test_y).float().mean().item()
            records.append({"train_loss": loss.item(), "test_acc": acc, "epoch": epoch + 1})
        with detach_interp_coeff(True):
            self._refresh_features(inputs, targets)
        self.eval()
        return records

    # ---- streaming
    def update(self, inputs, targets, update_stem=True, update_gp=True):
        inputs = inputs.view(-1, self.stem.input_dim)
        y, _, noise = dirichlet_transform(targets.view(-1), self.alpha_eps)
        stem_loss = self._update_stem(inputs, y, noise) if update_stem else 0.
        gp_loss = self._update_gp(inputs, y) if update_gp else 0.
        with torch.no_grad():
            self.gp.condition_on_observations(self.stem(inputs), y, noise, inplace=True)
            self._raw_inputs = [torch.cat([*self._raw_inputs, inputs])]
            if update_stem:
                self.stem.train()
                self._get_features(inputs)          # BatchNorm statistics
        self.eval()
        return stem_loss, gp_loss

    def _update_gp(self, inputs, targets):
        self.gp_optimizer.zero_grad()
        self.mll.train()
        self.gp.train()
        with settings.skip_logdet_forward(True):
            loss = -self.mll(self.gp(inputs), targets).sum()
        loss.backward()
        self.gp_optimizer.step()
        self.gp.zero_grad()
        self.gp.eval()
        return loss.item()

    def _update_stem(self, inputs, targets, noise):
        self.stem_optimizer.zero_grad()
        num_seen = self.gp.num_data
        new_features = self.stem(inputs)
        if new_features.requires_grad is False:
            return 0
        new_y = (targets / noise).t().unsqueeze(-1)                  # D^-1 y per output: [t, q, 1]
        loss = -sm_partial_mll(self.gp, new_features, new_y, num_seen).sum()
        loss.backward()
        self.stem_optimizer.step()
        return loss.item()

    def _get_features(self, inputs):
        """Stem forward on the new points plus a random replay minibatch (refreshes BatchNorm statistics)."""
        inputs = inputs.view(-1, self.stem.input_dim)
        seen = self._raw_inputs[0]
        replay = seen[torch.randint(0, seen.size(0), (1024,))]
        return self.stem(torch.cat([inputs, replay]))[:inputs.size(0)]

    def set_lr(self, gp_lr, stem_lr=None, bn_mom=None):
        self._make_optimizers(gp_lr, gp_lr if stem_lr is None else stem_lr)
        if bn_mom is not None:
            for mod in self.stem.modules():
                if isinstance(mod, torch.nn.BatchNorm1d):
                    mod.momentum = bn_mom
