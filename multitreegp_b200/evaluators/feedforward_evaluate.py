"""Drop-in for MultiTreeGP/evaluators/feedforward_evaluate.py (class ``Evaluator``, :10-116):
static symbolic state-feedback policies, u = trees(concat[obs, target]), evaluated by a fixed-step
ODE rollout of the environment; post-pass at the save points, env cost, non-finite -> 1e4, mean
over the batch, clip (:36-51).  Runs in the fused CUDA rollout kernel (kind MTGP_EVAL_STATIC).
"""
from __future__ import annotations

import numpy as np

from .. import abi
from ..rollout_data import RolloutData
from ._base import EvaluatorBase, _check_solver, _np


def control_rollout_data(self, data, state_size: int) -> RolloutData:
    x0, ts, targets, _process_keys, obs_keys, params = data  # feedforward_evaluate.py:62
    env = self.env
    params = np.stack([_np(p, np.float32).reshape(-1) for p in params], axis=1)  # pytree of [B] -> [B, n_params]
    if params.shape[1] != env.n_params:
        raise ValueError(f"{type(env).__name__} expects {env.n_params} parameter arrays, got {params.shape[1]}")
    keys = _np(obs_keys)
    if keys is not None:
        keys = keys.astype(np.uint32).reshape(-1, 2)
    return RolloutData(kind=self.kind, env_id=env.env_id, x0=_np(x0), ts=_np(ts), dt0=self.dt0,
                       max_fitness=self.max_fitness, targets=_np(targets), obs_keys=keys, params=params,
                       n_obs=env.n_obs, state_size=state_size, max_steps=self.max_steps,
                       obs_noise=float(env.obs_noise), threefry_partitionable=self.threefry_partitionable,
                       **self._solver_settings)


class Evaluator(EvaluatorBase):
    kind = abi.EVAL_STATIC

    def __init__(self, env, dt0: float, solver=None, max_steps: int = 16 ** 4, stepsize_controller=None,
                 **bind_kwargs) -> None:
        self._solver_settings = _check_solver(solver, stepsize_controller)
        super().__init__(**bind_kwargs)
        self.env = env
        self.max_fitness = 1e4
        self.obs_size = env.n_obs
        self.control_size = env.n_control
        self.latent_size = env.n_var * env.n_dim
        self.dt0 = dt0
        self.solver = solver
        self.max_steps = max_steps
        self.stepsize_controller = stepsize_controller

    def rollout_data(self, data) -> RolloutData:
        return control_rollout_data(self, data, 0)

    def evaluate_candidate(self, candidate, data, tree_evaluator=None):
        """Returns (xs, ys, us, fitness) like feedforward_evaluate.py:53-62."""
        out = self._trajectories(candidate, data)
        return out["xs"], out["ys"], out["us"], out["traj_fitness"]
