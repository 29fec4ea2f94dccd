"""Acrobot (reference: environments/control_environments/acrobot.py:7-87).

Device functor: ``Env<MTGP_ENV_ACROBOT>`` in csrc/mtgp_rollout.cu -- f_obs with the (-pi, pi] wrap
(:29-32), drift (:51-72), fitness_function (:77-84), cond_fn_nan (:86-87).
"""
from __future__ import annotations

import numpy as np

from ... import abi
from .control_environment_base import EnvironmentBase, as_rng


class Acrobot(EnvironmentBase):
    env_id = abi.ENV_ACROBOT
    n_params = 4

    def __init__(self, process_noise, obs_noise, n_obs=4):
        self.n_var = 4
        self.n_control = 1
        self.n_targets = 0
        self.n_dim = 1
        self.init_bounds = np.array([0.1, 0.1, 0.1, 0.1], dtype=np.float32)
        super().__init__(process_noise, obs_noise, self.n_var, self.n_control, self.n_dim, n_obs)

    def sample_init_states(self, batch_size, key):
        rng = as_rng(key)
        x0 = rng.uniform(-self.init_bounds, self.init_bounds, size=(batch_size, self.n_var)).astype(np.float32)
        targets = np.zeros((batch_size, self.n_targets), dtype=np.float32)
        return x0, targets

    def sample_params(self, batch_size, mode, ts, key):
        l1 = l2 = m1 = m2 = np.ones((batch_size,), dtype=np.float32)
        return l1, l2, m1, m2
