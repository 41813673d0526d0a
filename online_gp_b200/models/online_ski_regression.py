"""``OnlineSKIRegression`` — stem + WISKI GP + two Adam optimisers; ``update()`` = [stem SM-MLL step] + [GP
hyper-parameter step on the Woodbury MLL] + ``condition_on_observations`` in place.

Same class, constructor and methods as ``online_gp/models/online_ski_regression.py:15-198`` (``fit``, ``update``,
``predict``, ``evaluate``, ``set_lr``, ``set_train_data``, ``noise``); this is the object the streaming loop
``experiments/regression.py:48-54`` drives and the unit of work ``bench.py`` times.
"""
import torch
from torch.optim.lr_scheduler import CosineAnnealingLR

from .. import ops, settings
from ..graphs import StepGraphs, make_adam_capturable
from ..mlls.batched_woodbury_marginal_log_likelihood import BatchedWoodburyMarginalLogLikelihood
from ..mlls.streaming_partial_mll import sm_partial_mll
from ..settings import detach_interp_coeff
from .batched_fixed_noise_online_gp import FixedNoiseOnlineSKIGP


class OnlineSKIRegression(torch.nn.Module):
    def __init__(self, stem, init_x, init_y, lr, grid_size, grid_bound, covar_module=None, **kwargs):
        super().__init__()
        self.stem = stem.to(init_x.device)
        assert init_y.ndim == 2, "targets must have explicit output dimension"
        if init_y.size(-1) == 1:
            target_batch_shape = []
        else:
            target_batch_shape = torch.Size([init_y.size(-1)])
        features = self.stem(init_x).detach()
        noise_term = torch.ones_like(init_y)
        grid_bound += 1e-1
        self.gp = FixedNoiseOnlineSKIGP(
            features,
            init_y,
            noise_term,
            covar_module=covar_module,
            grid_bounds=torch.tensor([[-grid_bound, grid_bound]] * stem.output_dim),
            grid_size=[grid_size] * stem.output_dim,
            learn_additional_noise=True,
        )
        self.mll = BatchedWoodburyMarginalLogLikelihood(self.gp.likelihood, self.gp)
        self.gp_optimizer = torch.optim.Adam(self.gp.parameters(), lr=lr)
        self.stem_optimizer = torch.optim.Adam(self.stem.parameters(), lr=lr)
        self._target_batch_shape = target_batch_shape
        self.target_dim = init_y.size(-1)
        self._raw_inputs = [init_x]
        self._graphs = None          # opt-in CUDA-graph replay of evaluate() / update(): enable_cuda_graphs()

    def forward(self, inputs):
        inputs = inputs.view(-1, self.stem.input_dim)
        features = self.stem(inputs)
        return self.gp(features)

    def _reshape_targets(self, targets):
        targets = targets.view(-1, self.target_dim)
        if targets.size(-1) == 1:
            targets = targets.squeeze(-1)
        else:
            targets = targets.t()
        return targets

    def predict(self, inputs):
        self.eval()
        pred_dist = self(inputs)
        pred_mean = pred_dist.mean.reshape(-1, self.target_dim) if self.target_dim == 1 \
            else pred_dist.mean.t().reshape(-1, self.target_dim)
        pred_var = pred_dist.variance.reshape(-1, self.target_dim) if self.target_dim == 1 \
            else pred_dist.variance.t().reshape(-1, self.target_dim)
        pred_var = pred_var + self.gp.likelihood.second_noise.to(pred_var.dtype).reshape(1, -1)
        return pred_mean, pred_var

    def _evaluate_stats(self, input_batch, target_batch):
        """[rmse, nll] of one chunk as a device tensor (no host read)."""
        pred_mean, pred_var = self.predict(input_batch)
        rmse = (pred_mean - target_batch).pow(2).mean().sqrt()
        diag_dist = torch.distributions.Normal(pred_mean, pred_var.sqrt(), validate_args=False)
        nll = -diag_dist.log_prob(target_batch).mean()
        return torch.stack([rmse, nll])

    def evaluate(self, inputs, targets):
        inputs = inputs.view(-1, self.stem.input_dim)
        targets = targets.view(-1, self.target_dim)
        if self._graph_usable(inputs):
            return self._evaluate_graphed(inputs, targets)
        self._graph_phase(None)
        # Don't use `torch.no_grad` here, caches will be used for training
        self.eval()
        chunks = list(zip(inputs.split(1024), targets.split(1024)))
        # one device->host read for the whole call: the per-chunk statistics stay on the device and the interpolation
        # bounds flags are read back after them (the reference reads rmse and nll with one .item() each, :72-77)
        with settings.defer_interp_bounds_check(inputs.is_cuda):
            stats = [self._evaluate_stats(input_batch, target_batch) for input_batch, target_batch in chunks]
            rmse, nll = torch.stack(stats).mean(0).tolist()
        ops.flush_bounds_checks()
        return rmse, nll

    def fit(self, inputs, targets, num_epochs, test_dataset=None):
        self._graph_phase(None)
        records = []
        gp_lr_sched = CosineAnnealingLR(self.gp_optimizer, num_epochs, 1e-4)
        stem_lr_sched = CosineAnnealingLR(self.stem_optimizer, num_epochs, 1e-4)
        features = self._refresh_features(inputs, targets)
        for epoch in range(num_epochs):
            self.train()
            self.mll.train()
            self.stem_optimizer.zero_grad()
            self.gp_optimizer.zero_grad()
            train_dist = self.gp(features)
            loss = -self.mll(train_dist, targets).sum()
            loss.backward()
            self.stem_optimizer.step()
            self.gp_optimizer.step()
            stem_lr_sched.step()
            gp_lr_sched.step()
            features = self._refresh_features(inputs, targets)

            rmse = nll = float("NaN")
            if test_dataset is not None:
                test_x, test_y = test_dataset[:]
                rmse, nll = self.evaluate(test_x, test_y)
            records.append({"epoch": epoch + 1, "train_loss": loss.item(),
                            "test_rmse": rmse, "test_nll": nll,
                            "noise": self.gp.likelihood.second_noise_covar.noise.mean().item()})

        with detach_interp_coeff(True):
            self._refresh_features(inputs, targets)

        self.eval()
        return records

    def update(self, inputs, targets, update_stem=True, update_gp=True):
        inputs = inputs.view(-1, self.stem.input_dim)
        targets = targets.view(-1, self.target_dim)
        if update_gp and self._graph_usable(inputs) and self._graphs.phase == "evaluated":
            return self._update_graphed(inputs, targets)
        self._graph_phase(None)

        stem_loss = self._update_stem(inputs, targets) if update_stem else 0.
        gp_loss = self._update_gp_tensor(inputs, targets) if update_gp else 0.

        with torch.no_grad():
            features = self.stem(inputs)
            noise_term = torch.ones_like(targets)
            self.gp.condition_on_observations(features, targets, noise_term, inplace=True)
            self._raw_inputs = [torch.cat([*self._raw_inputs, inputs])]
            self.stem.train()
            if update_stem and self._stem_has_batchnorm():
                self._get_features(inputs)

        self.eval()
        # the loss is read back only now, with the conditioning kernels already queued behind the hyper-parameter step
        return stem_loss, (gp_loss.item() if torch.is_tensor(gp_loss) else gp_loss)

    def _stem_has_batchnorm(self):
        # `_get_features` only exists to refresh BatchNorm statistics (:133-134); without such layers it is a no-op
        return any(isinstance(m, torch.nn.modules.batchnorm._BatchNorm) for m in self.stem.modules())

    def _update_gp_tensor(self, inputs, targets):
        self.gp_optimizer.zero_grad()

        self.gp.train()
        self.mll.train()
        with settings.skip_logdet_forward(True):
            features = self.stem(inputs)
            train_dist = self.gp(features.detach())
            loss = -self.mll(train_dist, targets).sum()
        loss.backward()
        self.gp_optimizer.step()

        self.gp.zero_grad()
        self.gp.eval()
        return loss.detach()

    def _update_gp(self, inputs, targets):
        return self._update_gp_tensor(inputs, targets).item()

    def _update_stem(self, inputs, targets):
        self.stem_optimizer.zero_grad()
        num_seen = self.gp.num_data

        self.stem.eval()  # we want deterministic features, so BatchNorm should be in eval mode
        new_features = self.stem(inputs)
        if new_features.requires_grad is False:
            return 0

        targets = targets.transpose(-1, -2).unsqueeze(-1)
        loss = -sm_partial_mll(self.gp, new_features, targets, num_seen).sum()
        loss.backward()
        self.stem_optimizer.step()

        return loss.item()

    def _get_features(self, inputs):
        # update batch norm stats
        inputs = inputs.view(-1, self.stem.input_dim)
        num_inputs = inputs.size(0)
        num_seen = self._raw_inputs[0].size(0)
        batch_size = 1024
        batch_idxs = torch.randint(0, num_seen, (batch_size,))
        input_batch = self._raw_inputs[0][batch_idxs]
        input_batch = torch.cat([inputs, input_batch])
        features = self.stem(input_batch)
        return features[:num_inputs]

    def _refresh_features(self, inputs, targets):
        features = self.stem(inputs)
        self.set_train_data(features, targets)
        self.gp.zero_grad()  # dump W-related caches
        return features

    def set_train_data(self, inputs, targets):
        self._graph_phase(None)
        noise = torch.ones_like(targets)
        self.gp.set_train_data(inputs, targets, noise)

    def set_lr(self, gp_lr, stem_lr=None, bn_mom=None):
        stem_lr = gp_lr if stem_lr is None else stem_lr
        if self._graphs is not None:
            self.enable_cuda_graphs(True, warmup_calls=self._graphs.warm0)     # captured graphs hold the old optimiser
        self.gp_optimizer = torch.optim.Adam(self.gp.parameters(), lr=gp_lr)
        self.stem_optimizer = torch.optim.Adam(self.stem.parameters(), lr=stem_lr)
        if bn_mom is not None:
            for m in self.stem.modules():
                if isinstance(m, torch.nn.BatchNorm1d):
                    m.momentum = bn_mom

    @property
    def noise(self):
        return self.gp.likelihood.noise

    # ------------------------------------------------------------------ CUDA-graph replay (online_gp_b200/graphs.py)
    def enable_cuda_graphs(self, enabled=True, warmup_calls=2):
        """Opt in / out of replaying ``evaluate`` / ``update`` as captured CUDA graphs.  The first ``warmup_calls``
        evaluate/update pairs still run eagerly (library handles, lazy module loading), the next pair is captured,
        later pairs replay.  Calls that do not fit the captured form (different batch size, trainable stem,
        ``update()`` without a preceding ``evaluate()``) run eagerly on the same state."""
        self._graphs = StepGraphs(warmup_calls) if enabled else None
        return self

    def _graph_phase(self, phase):
        if self._graphs is not None:
            self._graphs.phase = phase

    def _graph_usable(self, inputs):
        G = self._graphs
        if G is None or G.failed or not inputs.is_cuda or inputs.shape[0] > 1024:
            return False
        if G.q is not None and inputs.shape[0] != G.q:
            return False
        if self._stem_has_batchnorm() or any(p.requires_grad for p in self.stem.parameters()):
            return False
        return True

    def _graph_fail(self, err):
        self._graphs.fail(err)
        self.gp._dump_caches()

    def _evaluate_graphed(self, inputs, targets):
        G = self._graphs
        if G.eval is None and G.warm > 0:
            G.warm -= 1
            self._graphs = None
            try:
                return self.evaluate(inputs, targets)
            finally:
                self._graphs = G
        if G.q is None:
            G.setup(inputs, targets)
            # the observation counter moves to the device so that the MLL's n-dependent terms follow the stream
            self.gp._num_data_t = torch.full((), float(self.gp.num_data), dtype=targets.dtype, device=inputs.device)
            make_adam_capturable(self.gp_optimizer)
        G.load(inputs, targets)
        if G.eval is None:
            self.eval()
            self.gp._dump_caches()
            try:
                G.eval = G.capture(lambda: self._evaluate_stats(G.x, G.y))
            except Exception as err:            # noqa: BLE001 - any capture problem means: run eagerly
                self._graph_fail(err)
                return self.evaluate(inputs, targets)
        rmse, nll = G.replay(G.eval)
        G.phase = "evaluated"
        return rmse, nll

    def _update_graphed(self, inputs, targets):
        G = self._graphs
        n_before = self.gp.num_data
        G.load(inputs, targets)
        if G.upd is None:
            def body():
                loss = self._update_gp_tensor(G.x, G.y)
                with torch.no_grad():
                    self.gp.condition_on_observations(self.stem(G.x), G.y, torch.ones_like(G.y), inplace=True)
                return loss
            try:
                G.upd = G.capture(body)
            except Exception as err:            # noqa: BLE001
                self.gp.num_data = n_before
                self._graph_fail(err)
                G.phase = None
                return self.update(inputs, targets)
        try:
            (gp_loss,) = G.replay(G.upd)
        finally:
            # also when the replay reports out-of-bounds inputs: the device-side state has already moved on (the
            # batch went in with clamped stencils), so the host-side bookkeeping must follow before the error surfaces
            G.phase = None
            self.gp.num_data = n_before + inputs.shape[0]
            self.gp._dump_caches()
            self._raw_inputs = [torch.cat([*self._raw_inputs, inputs])]
            self.eval()
        return 0., gp_loss

    @property
    def graph_launches(self):
        """Kernels of this library executed through graph replays so far (bench.py adds them to gpu_launches)."""
        return 0 if self._graphs is None else self._graphs.launches
