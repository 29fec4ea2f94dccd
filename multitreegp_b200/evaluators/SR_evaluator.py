"""Drop-in for MultiTreeGP/evaluators/SR_evaluator.py (class ``Evaluator``, :9-94).

The candidate's trees are the vector field of an ODE integrated from each x0 over ``ts`` with
fixed-step Euler; fitness = mean_k sum_j (pred - true)^2, non-finite -> 1e5, mean over the batch,
clip [0, 1e5] (:22-24, :42-45).  Runs in the fused CUDA rollout kernel (kind MTGP_EVAL_SR).
"""
from __future__ import annotations

import numpy as np

from .. import abi
from ..rollout_data import RolloutData
from ._base import EvaluatorBase, _check_solver, _np


class Evaluator(EvaluatorBase):
    kind = abi.EVAL_SR

    def __init__(self, solver=None, dt0: float = 0.01, max_steps: int = 16 ** 4, stepsize_controller=None,
                 **bind_kwargs) -> None:
        self._solver_settings = _check_solver(solver, stepsize_controller)
        super().__init__(**bind_kwargs)
        self.max_fitness = 1e5
        self.dt0 = dt0
        self.solver = solver
        self.stepsize_controller = stepsize_controller
        self.max_steps = max_steps

    def rollout_data(self, data) -> RolloutData:
        x0, ts, ys, _noise_keys = data  # (x0 [B,D], ts [K], ys [B,K,D], process_noise_keys) SR_evaluator.py:55
        return RolloutData(kind=self.kind, env_id=abi.ENV_NONE, x0=_np(x0), ts=_np(ts), dt0=self.dt0,
                           max_fitness=self.max_fitness, ys_true=_np(ys), max_steps=self.max_steps,
                           **self._solver_settings)

    def evaluate_candidate(self, candidate, data, tree_evaluator=None):
        """Returns (fitness [B], pred_ys [B,K,D]) like SR_evaluator.py:47-55."""
        out = self._trajectories(candidate, data)
        return out["traj_fitness"], out["xs"]
