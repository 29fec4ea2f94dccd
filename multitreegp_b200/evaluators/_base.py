"""Shared machinery of the three drop-in evaluators.

Reference protocol (kept): ``Evaluator.__call__(coefficients, nodes, data, tree_evaluator) -> fitness``
and ``evaluate_candidate(candidate, data, tree_evaluator)``; see SR_evaluator.py:30-55,
feedforward_evaluate.py:36-62, dynamic_evaluate.py:37-63.  ``tree_evaluator`` (the reference's
``GeneticProgramming.vmap_foriloop``) is accepted and ignored: trees are evaluated by the CUDA
interpreter.  Because the reference assigns node ids inside ``GeneticProgramming.__init__``
(genetic_programming.py:132-199), which the evaluator never sees, the evaluator must be bound to
the same ``operator_list`` / ``variable_list`` / ``layer_sizes`` (``bind`` or constructor kwargs).

Beyond the per-candidate protocol (which under jax.vmap becomes one FFI call per device shard),
``evaluate_population(flat_populations, data)`` is the coarse boundary that replaces
``strategy.jit_eval`` (genetic_programming.py:269,415) when JAX is not in the loop.
"""
from __future__ import annotations

import numpy as np

from .. import abi
from ..node_library import NodeLibrary
from ..rollout_data import RolloutData


def _np(a, dtype=None):
    if a is None:
        return None
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    a = np.asarray(a)
    return a.astype(dtype) if dtype is not None and a.dtype != dtype else a


def _check_solver(solver, stepsize_controller) -> dict:
    """Validates the solver / step-size controller (diffrax objects, multitreegp_b200.solvers stand-ins or
    names) and returns the matching RolloutData settings.  Euler + ConstantStepSize are the reference
    evaluators' class defaults; Dopri5 + PIDController is what the shipped notebooks pass."""
    from ..solvers import solver_settings
    return solver_settings(solver, stepsize_controller)


class EvaluatorBase:
    kind: int = -1
    max_fitness: float = 1e4

    def __init__(self, operator_list=None, variable_list=None, layer_sizes=None, library: NodeLibrary | None = None,
                 threefry_partitionable: bool = True) -> None:
        self.library = library
        self.threefry_partitionable = threefry_partitionable
        if library is None and operator_list is not None:
            self.bind(operator_list, variable_list, layer_sizes)
        self._engine = None
        self._dev_cache = (None, None)

    # -- node-table hand-off ---------------------------------------------------------------
    def bind(self, operator_list=None, variable_list=None, layer_sizes=None, library: NodeLibrary | None = None):
        """Rebuilds the reference's node-id assignment (genetic_programming.py:143-180)."""
        self.library = library if library is not None else NodeLibrary(operator_list, variable_list, layer_sizes)
        return self

    def _table(self) -> abi.NodeTable:
        if self.library is None:
            raise RuntimeError("evaluator is not bound to a node library: call bind(operator_list, variable_list, "
                               "layer_sizes) with the arguments given to GeneticProgramming")
        return self.library.table()

    def _get_engine(self):
        if self._engine is None:
            from ..engine import FitnessEngine  # imports torch; raises without CUDA
            self._engine = FitnessEngine()
        return self._engine

    def solver_settings(self) -> dict:
        """RolloutData fields of the evaluator's solver / step-size controller (solver, controller, rtol, atol,
        dtmin, pid, pid_limits): what the FFI binding forwards as attributes."""
        return dict(getattr(self, "_solver_settings", {}) or {})

    # -- to be provided by subclasses ------------------------------------------------------------
    def rollout_data(self, data) -> RolloutData:
        raise NotImplementedError

    # -- population-level boundary (replaces strategy.jit_eval) ---------------------------------
    def evaluate_population(self, flat_populations, data, size_parsimony: float = 0.0):
        """f32[P,T,N,4] -> f32[P].  numpy in -> numpy out (host path, copies inside the C call);
        CUDA torch tensor in -> CUDA torch tensor out (device path, asynchronous)."""
        rd = data if isinstance(data, RolloutData) else self.rollout_data(data)
        eng = self._get_engine()
        if isinstance(flat_populations, np.ndarray):
            pop = np.ascontiguousarray(flat_populations, dtype=np.float32)
            return eng.rollout_fitness_host(pop, self._table(), rd, size_parsimony)
        from ..engine import DeviceData
        key, dev = self._dev_cache
        if key is not rd:
            dev = DeviceData(rd, eng.device)
            self._dev_cache = (rd, dev)
        return eng.rollout_fitness(flat_populations.contiguous(), self._table(), dev, size_parsimony)["fitness"]

    # -- reference protocol -----------------------------------------------------------------
    def __call__(self, coefficients, nodes, data, tree_evaluator=None):
        """Fitness of one candidate (coefficients [T,N,1], nodes [T,N,3]) or, with leading batch
        dimensions, of a batch of candidates -- what jax.vmap turns the reference call into."""
        coefficients, nodes = _np(coefficients, np.float32), _np(nodes, np.float32)
        cand = np.concatenate([nodes, coefficients], axis=-1)
        lead = cand.shape[:-3]
        flat = np.ascontiguousarray(cand.reshape((-1,) + cand.shape[-3:]))
        fit = self.evaluate_population(flat, data)
        return fit.reshape(lead) if lead else np.float32(fit[0])

    def _trajectories(self, candidate, data):
        import torch
        from ..engine import DeviceData
        cand = np.ascontiguousarray(_np(candidate, np.float32))[None]
        rd = data if isinstance(data, RolloutData) else self.rollout_data(data)
        eng = self._get_engine()
        out = eng.rollout_fitness(torch.from_numpy(cand).to(eng.device), self._table(), DeviceData(rd, eng.device),
                                  0.0, trajectories=True)
        return {k: v[0].cpu().numpy() for k, v in out.items() if k != "fitness"}
