"""Parameter containers with the reference's names/shapes (src/model/layers/mlp.py:227-337); the
arithmetic runs in libgaot_b200.so through gaot_b200.ops."""
from __future__ import annotations

import torch
from torch import nn

from ... import ops


class LinearChannelMLP(nn.Module):
    """Linear stack with exact-erf GELU between layers (mlp.py:307-337); keys `fcs.{i}.{weight,bias}`."""

    def __init__(self, layers, non_linearity=None, dropout=0.0):
        super().__init__()
        self.n_layers = len(layers) - 1
        assert self.n_layers >= 1
        if dropout and dropout > 0.0:
            raise NotImplementedError("dropout > 0 is not part of the accelerated path")
        self.fcs = nn.ModuleList([nn.Linear(layers[j], layers[j + 1]) for j in range(self.n_layers)])

    def forward(self, x):
        shape = x.shape
        x = x.reshape(-1, shape[-1])
        for i, fc in enumerate(self.fcs):
            last = i == self.n_layers - 1
            # bf16 mode: hidden layers (K = hidden width) run on the tensor cores and hand bf16 to the next
            # layer; the first layer consumes raw coordinates and always stays fp32 (K = 2d is tiny anyway)
            x = ops.linear(x, fc.weight, fc.bias, act=ops.ACT_NONE if last else ops.ACT_GELU, tc=i > 0,
                           out_bf16=(i > 0 and not last))
        return x.view(*shape[:-1], x.shape[-1])


class ChannelMLP(nn.Module):
    """Stack of 1x1 Conv1d (mlp.py:227-305); with n_layers=1 a single affine map.  Unlike the
    reference this module takes channel-LAST input [B, rows, Cin] (the layout the reference permutes
    from and back to around every call, magno.py:273-274) and returns [B, rows, Cout]."""

    def __init__(self, in_channels, out_channels=None, hidden_channels=None, n_layers=2, n_dim=2,
                 non_linearity=None, dropout=0.0, **kwargs):
        super().__init__()
        self.n_layers = n_layers
        self.in_channels = in_channels
        self.out_channels = in_channels if out_channels is None else out_channels
        self.hidden_channels = in_channels if hidden_channels is None else hidden_channels
        if dropout and dropout > 0.0:
            raise NotImplementedError("dropout > 0 is not part of the accelerated path")
        self.fcs = nn.ModuleList()
        for i in range(n_layers):
            cin = self.in_channels if i == 0 else self.hidden_channels
            cout = self.out_channels if i == n_layers - 1 else self.hidden_channels
            self.fcs.append(nn.Conv1d(cin, cout, 1))

    def forward(self, x, add=None):
        shape = x.shape
        x = x.reshape(-1, shape[-1])
        for i, fc in enumerate(self.fcs):
            last = i == self.n_layers - 1
            x = ops.linear(x, fc.weight[:, :, 0], fc.bias, add=add if last else None,
                           act=ops.ACT_NONE if last else ops.ACT_GELU)
        return x.view(*shape[:-1], x.shape[-1])


class MLP(nn.Module):
    """Parameter container of reference mlp.py:43-72 as ConditionedNorm uses it (num_layers <= 2: one Linear, no activation)."""

    def __init__(self, input_size: int, output_size: int, hidden_size: int, num_layers: int = 3, activation: str = "swish"):
        super().__init__()
        if num_layers > 2:
            raise NotImplementedError("only the single-layer MLP of ConditionedNorm is on the accelerated path")
        self.layers = nn.ModuleList([nn.Linear(input_size, output_size)])

    def forward(self, x):
        return ops.linear(x, self.layers[-1].weight, self.layers[-1].bias)


class ConditionedNorm(nn.Module):
    """Time-conditioned affine correction (mlp.py:74-124): x * (1 + c * mlp_scale(c)) + c * mlp_bias(c), c [B,1]."""

    def __init__(self, input_size: int, output_size: int, hidden_size: int):
        super().__init__()
        self.mlp_scale = MLP(input_size, output_size, hidden_size, num_layers=2, activation="none")
        self.mlp_bias = MLP(input_size, output_size, hidden_size, num_layers=2, activation="none")
        for m in (self.mlp_scale, self.mlp_bias):
            for layer in m.layers:
                nn.init.normal_(layer.weight, std=0.01)

    def forward(self, c, x):
        if c is None:
            raise ValueError("use_conditional_norm=True needs a `condition` tensor [B, 1]")
        c = torch.as_tensor(c, dtype=torch.float32, device=x.device).reshape(x.shape[0], -1)
        scale = 1.0 + c * self.mlp_scale(c)            # [B, D]   (tiny: B x D)
        bias = c * self.mlp_bias(c)
        return ops.cond_affine(x, scale, bias)
