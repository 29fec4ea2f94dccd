"""Pointwise (X -> y) symbolic-regression evaluator for the scale-up config of BASELINE.json.

NOT in the reference: its "symbolic regression" integrates the trees as an ODE.  This evaluator is
defined with the reference's own primitives:
    pred    = vmap(tree_evaluator, in_axes=[None, 0])(candidate, X)      (genetic_programming.py:390-401)
    fitness = mean(sum(square(pred - Y), axis=-1)); non-finite -> 1e5; clip     (SR_evaluator.py:24,42-45)
and runs in the K1 interpreter kernel (csrc/mtgp_pointwise.cu).  ``data = (X [n, n_vars], Y [n, T])``.
"""
from __future__ import annotations

import numpy as np

from ._base import EvaluatorBase, _np


class Evaluator(EvaluatorBase):
    def __init__(self, **bind_kwargs) -> None:
        super().__init__(**bind_kwargs)
        self.max_fitness = 1e5

    def evaluate_population(self, flat_populations, data, size_parsimony: float = 0.0):
        eng = self._get_engine()
        X, Y = data
        if isinstance(flat_populations, np.ndarray):
            pop = np.ascontiguousarray(flat_populations, dtype=np.float32)
            X = np.ascontiguousarray(_np(X, np.float32))
            Y = np.ascontiguousarray(_np(Y, np.float32)).reshape(X.shape[0], -1)
            return eng.pointwise_fitness_host(pop, self._table(), X, Y, self.max_fitness, size_parsimony)
        return eng.pointwise_fitness(flat_populations.contiguous(), self._table(), X, Y, self.max_fitness,
                                     size_parsimony)

    def evaluate_candidate(self, candidate, data, tree_evaluator=None):
        """Returns (fitness, pred [n, T])."""
        import torch
        eng = self._get_engine()
        X, Y = data
        X = torch.as_tensor(_np(X, np.float32)).to(eng.device).contiguous()
        Y = torch.as_tensor(_np(Y, np.float32)).to(eng.device).reshape(X.shape[0], -1).contiguous()
        cand = torch.as_tensor(np.ascontiguousarray(_np(candidate, np.float32))[None]).to(eng.device)
        fit, pred = eng.pointwise_fitness(cand, self._table(), X, Y, self.max_fitness, predictions=True)
        return fit[0].item(), pred[0].cpu().numpy()
