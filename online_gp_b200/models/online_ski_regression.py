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
    """Public surface of the reference class: ``OnlineSKIRegression(stem, init_x, init_y, lr, grid_size, grid_bound,
    covar_module=None)`` with ``fit / update / predict / evaluate / set_lr / set_train_data / noise`` and the
    attributes ``gp``, ``mll``, ``stem``, ``gp_optimizer``, ``stem_optimizer``, ``target_dim``."""

    REPLAY_BATCH = 1024        # replay minibatch for BatchNorm statistics; also the chunk size of evaluate()

    def __init__(self, stem, init_x, init_y, lr, grid_size, grid_bound, covar_module=None, comm=None, **kwargs):
        """``comm`` (an ``online_gp_b200.parallel.Comm`` over a torch.distributed group, an addition to the reference
        signature): with more than one rank the inducing-grid rows are sharded across the ranks' GPUs and this object
        drives the row-sharded engine (``parallel.ShardedOnlineSKIRegression``: Identity stem, one output) through the
        same ``evaluate / update / predict / set_lr / noise`` calls — one class for N = 1 and N > 1."""
        super().__init__()
        assert init_y.ndim == 2, "targets must have explicit output dimension"
        self._engine = None
        if comm is not None and getattr(comm, "world", 1) > 1:
            self._init_sharded(stem, init_x, init_y, lr, grid_size, grid_bound, covar_module, comm)
            return
        n_out = init_y.size(-1)
        self.stem = stem.to(init_x.device)
        half_width = grid_bound + 1e-1                     # the grid reaches a little beyond the stated feature range
        dims = stem.output_dim
        self.gp = FixedNoiseOnlineSKIGP(
            self.stem(init_x).detach(), init_y, torch.ones_like(init_y),
            covar_module=covar_module,
            grid_bounds=torch.tensor([[-half_width, half_width]] * dims),
            grid_size=[grid_size] * dims,
            learn_additional_noise=True,
        )
        self.mll = BatchedWoodburyMarginalLogLikelihood(self.gp.likelihood, self.gp)
        self._new_optimizers(lr, lr)
        self.target_dim = n_out
        self._target_batch_shape = [] if n_out == 1 else torch.Size([n_out])
        self._raw_inputs = [init_x]
        self._graphs = None          # opt-in CUDA-graph replay of evaluate() / update(): enable_cuda_graphs()

    # ------------------------------------------------------------------ row-sharded engine (N > 1)
    def _init_sharded(self, stem, init_x, init_y, lr, grid_size, grid_bound, covar_module, comm):
        from ..parallel import ShardedOnlineSKIRegression
        if any(p.requires_grad for p in stem.parameters() if isinstance(p, torch.nn.Parameter)) or init_y.size(-1) != 1:
            raise NotImplementedError("the row-sharded engine serves an Identity stem and one output (SURVEY §8e)")
        self.stem = stem
        self.target_dim = 1
        self._graphs = None
        self._engine = ShardedOnlineSKIRegression(init_x, init_y, lr, grid_size, grid_bound, comm=comm, covar_module=covar_module)
        self.gp = self._engine            # covar_module, likelihood, num_data, panels (L_loc / B_loc) live on the engine

    @property
    def engine(self):
        """The row-sharded engine when this model was built with ``comm=`` over more than one rank, else None."""
        return self._engine

    def _new_optimizers(self, gp_lr, stem_lr):
        self.gp_optimizer = torch.optim.Adam(self.gp.parameters(), lr=gp_lr)
        self.stem_optimizer = torch.optim.Adam(self.stem.parameters(), lr=stem_lr)

    # ------------------------------------------------------------------ prediction
    def forward(self, inputs):
        return self.gp(self.stem(inputs.view(-1, self.stem.input_dim)))

    def _reshape_targets(self, targets):
        flat = targets.view(-1, self.target_dim)
        return flat.squeeze(-1) if self.target_dim == 1 else flat.t()

    def _columns(self, t):
        """[t, q] (or [q]) model output -> [q, target_dim]."""
        return t.reshape(-1, 1) if self.target_dim == 1 else t.t().reshape(-1, self.target_dim)

    def predict(self, inputs):
        """(mean, variance incl. the learned noise), each [q, target_dim]."""
        if self._engine is not None:
            return self._engine.predict(inputs.view(-1, self.stem.input_dim))
        self.eval()
        dist = self(inputs)
        var = self._columns(dist.variance)
        return self._columns(dist.mean), var + self.gp.likelihood.second_noise.to(var.dtype).reshape(1, -1)

    def _evaluate_stats(self, input_batch, target_batch):
        """[rmse, nll] of one chunk as a device tensor (no host read)."""
        mean, var = self.predict(input_batch)
        err = mean - target_batch
        gauss = torch.distributions.Normal(mean, var.sqrt(), validate_args=False)
        return torch.stack([err.pow(2).mean().sqrt(), -gauss.log_prob(target_batch).mean()])

    def evaluate(self, inputs, targets):
        """(rmse, nll) as Python floats.  Runs with autograd on: the caches built here are the ones the following
        ``update`` differentiates (``online_ski_regression.py:64-78``)."""
        inputs = inputs.view(-1, self.stem.input_dim)
        targets = targets.view(-1, self.target_dim)
        if self._engine is not None:
            return self._engine.evaluate(inputs, targets)
        if self._graph_usable(inputs):
            return self._evaluate_graphed(inputs, targets)
        self._graph_phase(None)
        self.eval()
        # one device->host read for the whole call: the per-chunk statistics stay on the device and the interpolation
        # bounds flags are read back after them (the reference reads rmse and nll with one .item() each)
        with settings.defer_interp_bounds_check(inputs.is_cuda):
            stats = [self._evaluate_stats(xb, yb)
                     for xb, yb in zip(inputs.split(self.REPLAY_BATCH), targets.split(self.REPLAY_BATCH))]
            rmse, nll = torch.stack(stats).mean(0).tolist()
        ops.flush_bounds_checks()
        return rmse, nll

    # ------------------------------------------------------------------ batch (pre-)training
    def fit(self, inputs, targets, num_epochs, test_dataset=None):
        """``num_epochs`` joint Adam steps on stem + GP hyper-parameters with cosine-annealed rates; the WISKI caches
        are rebuilt from the current features after every step (``:80-111``).  Returns one record per epoch."""
        if self._engine is not None:
            raise NotImplementedError("fit() is not available on the row-sharded engine")
        self._graph_phase(None)
        optimizers = (self.stem_optimizer, self.gp_optimizer)
        schedules = [CosineAnnealingLR(opt, num_epochs, 1e-4) for opt in optimizers]
        history = []
        features = self._refresh_features(inputs, targets)
        for epoch in range(1, num_epochs + 1):
            self.train()
            self.mll.train()
            for opt in optimizers:
                opt.zero_grad()
            loss = -self.mll(self.gp(features), targets).sum()
            loss.backward()
            for stepper in (*optimizers, *schedules):
                stepper.step()
            features = self._refresh_features(inputs, targets)
            rmse, nll = (float("NaN"), float("NaN")) if test_dataset is None else self.evaluate(*test_dataset[:])
            history.append({"epoch": epoch, "train_loss": loss.item(), "test_rmse": rmse, "test_nll": nll,
                            "noise": self.gp.likelihood.second_noise_covar.noise.mean().item()})
        with detach_interp_coeff(True):
            self._refresh_features(inputs, targets)
        self.eval()
        return history

    # ------------------------------------------------------------------ streaming
    def update(self, inputs, targets, update_stem=True, update_gp=True):
        """One streaming step (``:113-130``): stem step on the Sherman-Morrison partial MLL, hyper-parameter step on
        the Woodbury MLL, then the new batch is conditioned on in place.  Returns (stem_loss, gp_loss)."""
        inputs = inputs.view(-1, self.stem.input_dim)
        targets = targets.view(-1, self.target_dim)
        if self._engine is not None:
            return self._engine.update(inputs, targets)
        if update_gp and self._graph_usable(inputs) and self._graphs.phase == "evaluated":
            return self._update_graphed(inputs, targets)
        self._graph_phase(None)

        stem_loss = self._update_stem(inputs, targets) if update_stem else 0.
        self._prestart_condition(inputs, targets)
        gp_loss = self._update_gp_tensor(inputs, targets) if update_gp else 0.
        with torch.no_grad():
            self.gp.condition_on_observations(self.stem(inputs), targets, torch.ones_like(targets), inplace=True)
            self._raw_inputs = [torch.cat([*self._raw_inputs, inputs])]
            self.stem.train()
            if update_stem and self._stem_has_batchnorm():
                self._get_features(inputs)
        self.eval()
        # the loss is read back only now, with the conditioning kernels already queued behind the hyper-parameter step
        return stem_loss, (gp_loss.item() if torch.is_tensor(gp_loss) else gp_loss)

    def _stem_has_batchnorm(self):
        # `_get_features` only exists to refresh BatchNorm statistics; without such layers it is a no-op
        return any(isinstance(m, torch.nn.modules.batchnorm._BatchNorm) for m in self.stem.modules())

    def _prestart_condition(self, inputs, targets):
        """settings.overlap_root_update: the inverse-root half of the conditioning that ends this step goes to a side
        stream now, under the hyper-parameter step — only when the features do not depend on trainable stem parameters
        (then they are the same tensor values the conditioning will see after the step)."""
        if not (settings.overlap_root_update.on() and ops.overlap_capable(inputs)):
            return
        feats = self.stem(inputs)
        if feats.requires_grad:
            return
        with torch.no_grad():
            self.gp.prestart_condition(feats, targets, torch.ones_like(targets))

    def _update_gp_tensor(self, inputs, targets):
        """Adam step on -MLL (logdet value skipped, gradient kept); returns the loss as a detached device tensor."""
        self.gp_optimizer.zero_grad()
        self.gp.train()
        self.mll.train()
        with settings.skip_logdet_forward(True):
            dummy = self.gp(self.stem(inputs).detach())
            loss = -self.mll(dummy, targets).sum()
        loss.backward()
        self.gp_optimizer.step()
        self.gp.zero_grad()              # also drops the caches that depended on the old hyper-parameters
        self.gp.eval()
        return loss.detach()

    def _update_gp(self, inputs, targets):
        return self._update_gp_tensor(inputs, targets).item()

    def _update_stem(self, inputs, targets):
        """Stem step on the one-point Sherman-Morrison MLL increment (``:148-162``); 0 for a stem without parameters."""
        self.stem_optimizer.zero_grad()
        seen = self.gp.num_data
        self.stem.eval()                 # deterministic features: BatchNorm uses its running statistics
        feats = self.stem(inputs)
        if feats.requires_grad is False:
            return 0
        loss = -sm_partial_mll(self.gp, feats, targets.transpose(-1, -2).unsqueeze(-1), seen).sum()
        loss.backward()
        self.stem_optimizer.step()
        return loss.item()

    def _get_features(self, inputs):
        """Stem forward on the new points plus a random replay minibatch (refreshes BatchNorm statistics)."""
        inputs = inputs.view(-1, self.stem.input_dim)
        seen = self._raw_inputs[0]
        replay = seen[torch.randint(0, seen.size(0), (self.REPLAY_BATCH,))]
        return self.stem(torch.cat([inputs, replay]))[:inputs.size(0)]

    def _refresh_features(self, inputs, targets):
        feats = self.stem(inputs)
        self.set_train_data(feats, targets)
        self.gp.zero_grad()              # dump the caches built on the previous features
        return feats

    def set_train_data(self, inputs, targets):
        self._graph_phase(None)
        self.gp.set_train_data(inputs, targets, torch.ones_like(targets))

    def set_lr(self, gp_lr, stem_lr=None, bn_mom=None):
        if self._engine is not None:
            if self._engine._graphs is not None:
                self._engine.enable_cuda_graphs(True, warmup_calls=self._engine._graphs.warm0)
            self._engine.gp_optimizer = torch.optim.Adam(self._engine.parameters(), lr=gp_lr)
            return
        if self._graphs is not None:
            self.enable_cuda_graphs(True, warmup_calls=self._graphs.warm0)     # captured graphs hold the old optimiser
        self._new_optimizers(gp_lr, gp_lr if stem_lr is None else stem_lr)
        if bn_mom is not None:
            for mod in self.stem.modules():
                if isinstance(mod, torch.nn.BatchNorm1d):
                    mod.momentum = bn_mom

    @property
    def noise(self):
        if self._engine is not None:
            return self._engine._noise().reshape(1)
        return self.gp.likelihood.noise

    # ------------------------------------------------------------------ CUDA-graph replay (online_gp_b200/graphs.py)
    def enable_cuda_graphs(self, enabled=True, warmup_calls=2):
        """Opt in / out of replaying ``evaluate`` / ``update`` as captured CUDA graphs.  The first ``warmup_calls``
        evaluate/update pairs still run eagerly (library handles, lazy module loading), the next pair is captured,
        later pairs replay.  Calls that do not fit the captured form (different batch size, trainable stem,
        ``update()`` without a preceding ``evaluate()``) run eagerly on the same state."""
        if self._engine is not None:
            self._engine.enable_cuda_graphs(enabled, warmup_calls)
            return self
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
                self._prestart_condition(G.x, G.y)
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
        if self._engine is not None:
            return self._engine.graph_launches
        return 0 if self._graphs is None else self._graphs.launches
