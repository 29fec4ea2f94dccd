"""StirredTankReactor (reference: environments/control_environments/reactor.py:7-99).

Device functor: ``Env<MTGP_ENV_REACTOR>`` -- drift with the Arrhenius term (:78-87), cost (:92-96),
cond_fn_nan (:98-99), observation noise scaled by [15, 15, 0.1] (:43).
"""
from __future__ import annotations

import numpy as np

from ... import abi
from .control_environment_base import EnvironmentBase, as_rng


class StirredTankReactor(EnvironmentBase):
    env_id = abi.ENV_REACTOR
    n_params = 8

    def __init__(self, process_noise, obs_noise, n_obs=3, n_targets=1):
        self.n_var = 3
        self.n_control = 1
        self.n_dim = 1
        self.n_targets = n_targets
        self.init_lower_bounds = np.array([275, 350, 0.5], dtype=np.float32)
        self.init_upper_bounds = np.array([300, 375, 1.0], dtype=np.float32)
        super().__init__(process_noise, obs_noise, self.n_var, self.n_control, self.n_dim, n_obs)

    def sample_params(self, batch_size, mode, ts, key):
        rng = as_rng(key)
        ones = np.ones(batch_size, dtype=np.float32)
        if mode == "Constant":
            vals = [100 * ones, 239 * ones, -5.0e4 * ones, 5.0e4 * ones, 100 * ones, 300 * ones, 300 * ones, 20.0 * ones]
        elif mode == "Different":
            bounds = [(75, 150), (200, 350), (-55000, -45000), (25000, 75000), (75, 125), (300, 350), (250, 300),
                      (10, 30)]
            vals = [rng.uniform(lo, hi, size=(batch_size,)) for lo, hi in bounds]
        else:
            raise NotImplementedError(mode)
        return tuple(np.asarray(v, dtype=np.float32) for v in vals)

    def sample_init_states(self, batch_size, key):
        rng = as_rng(key)
        x0 = rng.uniform(self.init_lower_bounds, self.init_upper_bounds, size=(batch_size, self.n_var))
        targets = rng.uniform(400, 500, size=(batch_size, self.n_targets))
        return x0.astype(np.float32), targets.astype(np.float32)
