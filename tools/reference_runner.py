"""Runs the REAL reference (sdevries0/MultiTreeGP: jax + diffrax) on a case built by this repository's
test builders.  Importable only where jax, diffrax and the reference sources are available; used by
tools/record_reference_golden.py (golden vectors) and by `bench.py --impl reference` (CPU timing).

Nothing here restates the reference: the strategy, the evaluators and the environments are the
reference's own classes (MultiTreeGP/genetic_programming.py, evaluators/*.py, environments/**), called
through their public API: GeneticProgramming(...).jit_eval(flat_population, data) -- the hot entry point,
genetic_programming.py:249-269, :415 -- and Evaluator.evaluate_candidate (plot cells of the notebooks).
"""
from __future__ import annotations

import importlib
import os
import sys

import numpy as np


def find_reference() -> str | None:
    """Directory that contains the `MultiTreeGP` package: $MTGP_REFERENCE, baseline/_ref, /root/reference."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for cand in (os.environ.get("MTGP_REFERENCE"), os.path.join(root, "baseline", "_ref"), "/root/reference"):
        if cand and os.path.isdir(os.path.join(cand, "MultiTreeGP")):
            return cand
    return None


def available() -> tuple[bool, str]:
    """(True, versions) when the real reference can run here, else (False, why)."""
    ref = find_reference()
    if ref is None:
        return False, "reference sources not found (set MTGP_REFERENCE)"
    missing = []
    for mod in ("jax", "diffrax", "optax", "sympy"):
        try:
            importlib.import_module(mod)
        except Exception as e:  # noqa: BLE001
            missing.append(f"{mod}: {type(e).__name__}")
    if missing:
        return False, "cannot import " + ", ".join(missing)
    return True, versions()


def versions() -> str:
    out = []
    for mod in ("jax", "jaxlib", "diffrax", "equinox", "optax", "numpy"):
        try:
            out.append(f"{mod} {importlib.import_module(mod).__version__}")
        except Exception:  # noqa: BLE001
            out.append(f"{mod} ?")
    return ", ".join(out)


def _jnp_operator(string: str):
    """The node functions of the notebooks (examples/*.ipynb operator lists) plus this repository's
    extension set, as the callables the reference expects: (string, callable, arity, probability)."""
    import jax.numpy as jnp
    table = {
        "+": (lambda x, y: x + y, 2), "-": (lambda x, y: x - y, 2), "*": (lambda x, y: x * y, 2),
        "/": (lambda x, y: x / y, 2), "max": (lambda x, y: jnp.maximum(x, y), 2),
        "min": (lambda x, y: jnp.minimum(x, y), 2), "sin": (lambda x: jnp.sin(x), 1), "cos": (lambda x: jnp.cos(x), 1),
        "exp": (lambda x: jnp.exp(x), 1), "log": (lambda x: jnp.log(x), 1), "sqrt": (lambda x: jnp.sqrt(x), 1),
        "tanh": (lambda x: jnp.tanh(x), 1), "neg": (lambda x: -x, 1), "abs": (lambda x: jnp.abs(x), 1),
        "square": (lambda x: x * x, 1),
    }
    return table[string]


class ReferenceCase:
    """One (node library, population, RolloutData) case bound to the reference's own objects."""

    def __init__(self, lib, operator_list, variable_list, layer_sizes, pop: np.ndarray, data, env_name: str | None,
                 solver: str = "euler", threefry_partitionable: bool = True) -> None:
        ref = find_reference()
        if ref not in sys.path:
            sys.path.insert(0, ref)
        import jax
        jax.config.update("jax_threefry_partitionable", bool(threefry_partitionable))
        import diffrax
        import jax.numpy as jnp
        from MultiTreeGP.genetic_programming import GeneticProgramming
        from multitreegp_b200 import abi

        self.jnp, self.jax = jnp, jax
        P, T, N, _ = pop.shape
        if data.solver == abi.SOLVER_DOPRI5:
            solver_obj = diffrax.Dopri5()
        else:
            solver_obj = diffrax.Euler()
        if data.controller == abi.CONTROLLER_PID:
            kw = dict(rtol=float(data.rtol), atol=float(data.atol))
            if data.dtmin > 0:
                kw.update(dtmin=float(data.dtmin), force_dtmin=True)
            controller = diffrax.PIDController(**kw)
        else:
            controller = diffrax.ConstantStepSize()
        if data.kind == abi.EVAL_SR:
            from MultiTreeGP.evaluators.SR_evaluator import Evaluator
            self.evaluator = Evaluator(solver=solver_obj, dt0=float(data.dt0), max_steps=int(data.max_steps),
                                       stepsize_controller=controller)
            B = data.x0.shape[0]
            keys = jnp.zeros((B, 2), jnp.uint32)  # process-noise keys: a dead input (SR_evaluator.py:57-83)
            self.data = (jnp.asarray(data.x0), jnp.asarray(data.ts), jnp.asarray(data.ys_true), keys)
        else:
            env_mod = {"acrobot": ("acrobot", "Acrobot"), "harmonic": ("harmonic_oscillator", "HarmonicOscillator"),
                       "reactor": ("reactor", "StirredTankReactor")}[env_name]
            cls = getattr(importlib.import_module("MultiTreeGP.environments.control_environments." + env_mod[0]), env_mod[1])
            env = cls(0.0, float(data.obs_noise), n_obs=int(data.n_obs))
            if data.kind == abi.EVAL_STATIC:
                from MultiTreeGP.evaluators.feedforward_evaluate import Evaluator
                self.evaluator = Evaluator(env, float(data.dt0), solver=solver_obj, max_steps=int(data.max_steps),
                                           stepsize_controller=controller)
            else:
                from MultiTreeGP.evaluators.dynamic_evaluate import Evaluator
                self.evaluator = Evaluator(env, int(data.state_size), float(data.dt0), solver=solver_obj,
                                           max_steps=int(data.max_steps), stepsize_controller=controller)
            keys = jnp.asarray(data.obs_keys, jnp.uint32)
            params = tuple(jnp.asarray(data.params[:, i]) for i in range(data.params.shape[1]))
            self.data = (jnp.asarray(data.x0), jnp.asarray(data.ts), jnp.asarray(data.targets), jnp.zeros_like(keys), keys,
                         params)
        ops = []
        for tup in operator_list:
            string, prob = tup[0], (tup[-1] if len(tup) >= 3 else 1.0)
            fn, arity = _jnp_operator(string)
            ops.append((string, fn, arity, float(prob)))
        pop_size = P + (P % 2)  # the constructor wants an even population size; evaluation takes any P
        self.strategy = GeneticProgramming(1, pop_size, self.evaluator, ops, [list(v) for v in variable_list],
                                           jnp.asarray(layer_sizes), num_populations=1, max_nodes=int(N))
        # the id assignment must be the one the population was built with (genetic_programming.py:143-180)
        for s, i in lib.string_to_node.items():
            assert int(self.strategy.string_to_node[s]) == int(i), (s, i, self.strategy.string_to_node[s])
        self.pop = jnp.asarray(pop)

    def fitness(self, start: int = 0, stop: int | None = None) -> np.ndarray:
        """strategy.jit_eval on a slice of the flattened population (genetic_programming.py:415)."""
        out = self.strategy.jit_eval(self.pop[start:stop], self.data)
        return np.asarray(self.jax.block_until_ready(out), np.float32)

    def candidate(self, index: int) -> dict:
        """Evaluator.evaluate_candidate for one individual (trajectories and per-trajectory fitness)."""
        from multitreegp_b200 import abi
        out = self.evaluator.evaluate_candidate(self.pop[index], self.data, self.strategy.vmap_foriloop)
        out = [np.asarray(o, np.float32) for o in out]
        if len(out) == 2:      # SR: (fitness, pred_ys)
            return dict(traj_fitness=out[0], xs=out[1])
        if len(out) == 4:      # static: (xs, ys, us, fitness)
            return dict(xs=out[0], ys=out[1], us=out[2], traj_fitness=out[3])
        return dict(xs=out[0], ys=out[1], us=out[2], activities=out[3], traj_fitness=out[4])
