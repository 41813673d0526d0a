"""Feature stems (``online_gp/models/stems.py:4-62``): plain torch.nn modules in front of the GP, not a kernel
target; same class names, constructor arguments and forward semantics (features squashed into (-1, 1)
by ``tanh(z / 2)`` so that they stay inside the inducing grid)."""
import torch
from torch import nn


class Identity(nn.Module):
    """Pass-through stem.  ``parameters()`` hands the optimiser one dummy tensor (an optimiser refuses an empty
    parameter list) and ``modules()`` is empty, as in the reference."""

    def __init__(self, input_dim):
        super().__init__()
        self.input_dim = self.output_dim = input_dim

    def forward(self, inputs):
        return inputs

    def parameters(self, **kwargs):
        return [torch.eye(self.input_dim)]

    def modules(self):
        return []


def _squash(z):
    return torch.tanh(z / 2)


class LinearStem(nn.Sequential):
    """Linear -> BatchNorm (no affine) -> tanh(z / 2)."""

    def __init__(self, input_dim, feature_dim):
        super().__init__(nn.Linear(input_dim, feature_dim), nn.BatchNorm1d(feature_dim, affine=False))
        self.input_dim, self.output_dim = input_dim, feature_dim

    def forward(self, input):
        return _squash(super().forward(input))


class MLP(nn.Sequential):
    """``depth`` ReLU layers of widths ``hidden_dims`` (list or comma-separated string), then Linear -> BatchNorm
    (no affine, momentum 0.1) -> tanh(z / 2)."""

    def __init__(self, input_dim, feature_dim, depth, hidden_dims):
        widths = [int(w) for w in hidden_dims.split(",")] if isinstance(hidden_dims, str) else list(hidden_dims)
        layers, fan_in = [], input_dim
        for k in range(depth):
            layers += [nn.Linear(fan_in, widths[k]), nn.ReLU()]
            fan_in = widths[k]
        layers += [nn.Linear(widths[-1], feature_dim), nn.BatchNorm1d(feature_dim, affine=False, momentum=1e-1)]
        super().__init__(*layers)
        self.input_dim, self.output_dim = input_dim, feature_dim

    def forward(self, input):
        return _squash(super().forward(input))
