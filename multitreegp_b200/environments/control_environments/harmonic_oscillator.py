"""HarmonicOscillator (reference: environments/control_environments/harmonic_oscillator.py:8-84).

Device functor: ``Env<MTGP_ENV_HARMONIC>`` -- drift A@x + b@u (:70-71), quadratic cost (:76-81),
cond_fn_nan (:83-84).  Only per-trajectory constant (omega, zeta) are supported ("Constant" and
"Different" modes); the time-varying "Switch"/"Decay" modes belong to ChangingHarmonicOscillator,
which the shipped evaluators cannot run (no cond_fn_nan) and which is out of scope.
"""
from __future__ import annotations

import numpy as np

from ... import abi
from .control_environment_base import EnvironmentBase, as_rng


class HarmonicOscillator(EnvironmentBase):
    env_id = abi.ENV_HARMONIC
    n_params = 2

    def __init__(self, process_noise, obs_noise, n_obs=2):
        self.n_dim = 1
        self.n_var = 2
        self.n_control = 1
        self.n_targets = 1
        self.mu0 = np.zeros(self.n_var, dtype=np.float32)
        self.P0 = np.eye(self.n_var, dtype=np.float32) * np.array([3.0, 1.0], dtype=np.float32)
        super().__init__(process_noise, obs_noise, self.n_var, self.n_control, self.n_dim, n_obs)
        self.q = self.r = 0.5

    def sample_init_states(self, batch_size, key):
        rng = as_rng(key)
        x0 = self.mu0 + rng.standard_normal((batch_size, self.n_var)).astype(np.float32) @ self.P0
        targets = rng.uniform(-3, 3, size=(batch_size, self.n_targets)).astype(np.float32)
        return x0.astype(np.float32), targets

    def sample_params(self, batch_size, mode, ts, key):
        rng = as_rng(key)
        if mode == "Constant":
            omegas = np.ones((batch_size,), dtype=np.float32)
            zetas = np.zeros((batch_size,), dtype=np.float32)
        elif mode == "Different":
            omegas = rng.uniform(0.0, 2.0, size=(batch_size,)).astype(np.float32)
            zetas = rng.uniform(0.0, 1.5, size=(batch_size,)).astype(np.float32)
        else:
            raise NotImplementedError(f"param mode {mode!r}: time-varying parameters are out of scope (SURVEY.md 8f)")
        return omegas, zetas
