"""Device-side `DynamicPairDataset` (reference src/datasets/data_utils.py:73-235): the time-pair samples of sequential
training, generated where the data already lives.

Same constructor arguments, the same pair enumeration (`_generate_time_pairs`, data_utils.py:121-160), the same
`__len__` / `__getitem__` contract (input [num_nodes, u + c + 2] = [u_norm | c_norm | start_time | time_diff], target by
stepper mode, plus the coordinates in vx mode).  B200-native addition: `get_batch(indices)` assembles a whole batch with a
handful of batched gathers on the device `u_data` lives on (the reference builds every sample on the host in a DataLoader
worker and the trainer moves each batch with a blocking .to(device)), so a captured training step can be fed without a
host round trip.  Integer pair bookkeeping is numpy exactly as in the reference; arithmetic is torch on the data's device."""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch
from torch.utils.data import Dataset


class DynamicPairDataset(Dataset):
    def __init__(self, u_data: torch.Tensor, c_data: Optional[torch.Tensor], t_values: torch.Tensor, metadata,
                 max_time_diff: int = 14, time_step: int = 2, stepper_mode: str = "output", stats: Optional[dict] = None,
                 use_time_norm: bool = True, dataset_name: Optional[str] = None, x_data: Optional[torch.Tensor] = None,
                 is_variable_coords: bool = False):
        if stepper_mode not in ("output", "residual", "time_der"):
            raise ValueError(f"Unsupported stepper_mode: {stepper_mode}")
        self.dataset_name = dataset_name
        self.u_data, self.c_data, self.x_data = u_data, c_data, x_data
        self.t_values = torch.as_tensor(t_values)
        self.metadata = metadata
        self.stepper_mode = stepper_mode
        self.stats = stats
        self.use_time_norm = use_time_norm
        self.is_variable_coords = is_variable_coords
        self.num_samples, self.num_timesteps, self.num_nodes, self.num_vars = u_data.shape
        self.num_timesteps = min(self.num_timesteps - 1, max_time_diff)
        self.t_values = self.t_values[:self.num_timesteps + 1]
        self._generate_time_pairs(self.num_timesteps, time_step)

    def _generate_time_pairs(self, num_timesteps: int, time_step: int):
        t_in, t_out = [], []
        for lag in range(time_step, num_timesteps + 1, time_step):
            for i in range(0, num_timesteps - lag + 1, time_step):
                t_in.append(i)
                t_out.append(i + lag)
        self.t_in_indices = np.array(t_in)
        self.t_out_indices = np.array(t_out)
        tv = self.t_values.detach().cpu()
        self.time_diffs = tv[self.t_out_indices] - tv[self.t_in_indices]
        if self.use_time_norm and self.stats is not None:
            self.start_times = tv[self.t_in_indices]
            self.start_times_norm = (self.start_times - self.stats["start_time"]["mean"]) / self.stats["start_time"]["std"]
            self.time_diffs_norm = (self.time_diffs - self.stats["time_diffs"]["mean"]) / self.stats["time_diffs"]["std"]
        else:
            self.start_times_norm = tv[self.t_in_indices]
            self.time_diffs_norm = self.time_diffs
        dev, dt = self.u_data.device, self.u_data.dtype
        # per-pair scalars on the data's device (the reference expands them to [pairs, num_nodes, 1] on the host)
        self._t0 = torch.as_tensor(self.start_times_norm).to(dev, dt)
        self._dt = torch.as_tensor(self.time_diffs_norm).to(dev, dt)
        self._dt_actual = torch.as_tensor(self.time_diffs).to(dev, dt)
        self._tin = torch.as_tensor(self.t_in_indices, device=dev)
        self._tout = torch.as_tensor(self.t_out_indices, device=dev)

    def __len__(self):
        return self.num_samples * len(self.t_in_indices)

    def _stat(self, key, which):
        return torch.as_tensor(self.stats[key][which]).to(self.u_data.device, self.u_data.dtype)

    def get_batch(self, indices):
        """Samples `indices` (any integer sequence / tensor) as one batch: (input [B, N, u+c+2], target [B, N, u][, x [B, N, d]])."""
        dev = self.u_data.device
        idx = torch.as_tensor(indices, device=dev, dtype=torch.int64)
        n_pairs = len(self.t_in_indices)
        sample, pair = idx // n_pairs, idx % n_pairs
        tin, tout = self._tin[pair], self._tout[pair]
        u_in, u_out = self.u_data[sample, tin], self.u_data[sample, tout]           # [B, N, U]
        u_in_norm = (u_in - self._stat("u", "mean")) / self._stat("u", "std") if self.stats is not None else u_in
        feats = [u_in_norm]
        if self.c_data is not None:
            c_in = self.c_data[sample, tin]
            if self.stats is not None and "c" in self.stats:
                c_in = (c_in - self._stat("c", "mean")) / self._stat("c", "std")
            feats.append(c_in)
        B, N = u_in.shape[0], u_in.shape[1]
        feats.append(self._t0[pair].view(B, 1, 1).expand(B, N, 1))
        feats.append(self._dt[pair].view(B, 1, 1).expand(B, N, 1))
        inp = torch.cat(feats, dim=-1)
        if self.stepper_mode == "output":
            target = (u_out - self._stat("u", "mean")) / self._stat("u", "std")
        elif self.stepper_mode == "residual":
            target = u_out - u_in
            if self.stats is not None:
                target = (target - self._stat("res", "mean")) / self._stat("res", "std")
        else:
            target = (u_out - u_in) / self._dt_actual[pair].view(B, 1, 1)
            if self.stats is not None:
                target = (target - self._stat("der", "mean")) / self._stat("der", "std")
        if self.is_variable_coords and self.x_data is not None:
            return inp, target, self.x_data[sample, tin]
        return inp, target

    def __getitem__(self, idx):
        out = self.get_batch([int(idx)])
        return tuple(t[0] for t in out)
