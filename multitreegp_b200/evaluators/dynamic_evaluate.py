"""Drop-in for MultiTreeGP/evaluators/dynamic_evaluate.py (class ``Evaluator``, :10-125): dynamic
symbolic policies -- a latent state a (``state_size`` trees, da/dt = trees(concat[y, a, u, target]))
and a readout (u = trees(concat[0, a, 0, target])) integrated together with the environment.
Runs in the fused CUDA rollout kernel (kind MTGP_EVAL_DYNAMIC).  The legacy ``EvaluatorMT``
(:128-217) is dead code in the reference and is not provided.
"""
from __future__ import annotations

from .. import abi
from ..rollout_data import RolloutData
from ._base import EvaluatorBase, _check_solver
from .feedforward_evaluate import control_rollout_data


class Evaluator(EvaluatorBase):
    kind = abi.EVAL_DYNAMIC

    def __init__(self, env, state_size: int, dt0: float, solver=None, max_steps: int = 16 ** 4,
                 stepsize_controller=None, **bind_kwargs) -> None:
        self._solver_settings = _check_solver(solver, stepsize_controller)
        super().__init__(**bind_kwargs)
        self.env = env
        self.max_fitness = 1e4
        self.state_size = state_size
        self.obs_size = env.n_obs
        self.control_size = env.n_control
        self.latent_size = env.n_var * env.n_dim
        self.dt0 = dt0
        self.solver = solver
        self.max_steps = max_steps
        self.stepsize_controller = stepsize_controller

    def rollout_data(self, data) -> RolloutData:
        return control_rollout_data(self, data, self.state_size)

    def evaluate_candidate(self, candidate, data, tree_evaluator=None):
        """Returns (xs, ys, us, activities, fitness) like dynamic_evaluate.py:54-63."""
        out = self._trajectories(candidate, data)
        return out["xs"], out["ys"], out["us"], out["activities"], out["traj_fitness"]
