"""Host-side mirror of the reference's EnvironmentBase (control_environment_base.py:22-63).

The reference environments are JAX objects whose methods (f_obs, drift, fitness_function,
cond_fn_nan) run inside the traced evaluator.  Here those methods live on the GPU as compiled-in
device functors of the rollout kernel (multitreegp_b200/csrc/mtgp_rollout.cu, ``Env<ENV>``); the
Python class only carries what the evaluator needs to build the kernel descriptor (``env_id``,
dimensions, noise levels) plus NumPy versions of the data-preparation helpers
(``sample_init_states``, ``sample_params``) the notebooks call before the generation loop.
``key`` arguments accept an int seed or a numpy Generator (jax.random keys do not exist here).
"""
from __future__ import annotations

import abc

import numpy as np


def as_rng(key) -> np.random.Generator:
    if isinstance(key, np.random.Generator):
        return key
    return np.random.default_rng(key)


class EnvironmentBase(abc.ABC):
    env_id: int = 0
    n_params: int = 0

    def __init__(self, process_noise, obs_noise, n_var, n_control, n_dim, n_obs):
        self.process_noise = process_noise  # dead input of the shipped evaluators (only the drift is solved)
        self.obs_noise = obs_noise
        self.n_var = n_var
        self.n_control = n_control
        self.n_dim = n_dim
        self.n_obs = n_obs

    @abc.abstractmethod
    def sample_init_states(self, batch_size, key):
        return

    @abc.abstractmethod
    def sample_params(self, batch_size, mode, ts, key):
        return
