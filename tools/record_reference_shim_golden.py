"""Records golden vectors from the reference's OWN Python code, executed in this image over a numpy stand-in for jax.

    python tools/record_reference_shim_golden.py        ->  tests/golden/refshim/refshim_env.npz, refshim_trees.npz

jax / diffrax are not installable here, so the real reference cannot run end to end (tools/record_reference_golden.py
is the recorder for a machine that has them).  What CAN run here is every line of the reference that is not
third-party: tools/refshim/ provides `jax.numpy` on numpy float32 (weak python scalars, `.at[]` updates), `jax.lax`
control flow as python control flow (switch clamps its index), `jax.vmap` as a loop, and inert stubs for jit / sharding /
diffrax / optax / jax.random.  With it this script imports /root/reference/MultiTreeGP unmodified and records

  refshim_trees.npz   GeneticProgramming.__init__'s node library (string_to_node: genetic_programming.py:132-199) and
                      GeneticProgramming.vmap_foriloop (body_fun / foriloop, :356-401) on sampled and hand-made trees
  refshim_env.npz     Acrobot / HarmonicOscillator / StirredTankReactor: drift, f_obs (noise term = 0: jax.random is
                      third-party), cond_fn_nan, fitness_function (acrobot.py:29-87, harmonic_oscillator.py:58-84,
                      reactor.py:22-99, control_environment_base.py:43-48)

tests/test_reference_shim_golden.py checks the CPU oracle against both files.  + - * / are IEEE-exact in numpy as in
XLA, so trees built from them are compared bit for bit; sin / cos / exp go through numpy's float32 routines instead of
XLA's (both within a few ulp of the contract's), so those are compared with rtol 1e-5 -- the north star's tolerance.
Not covered, and still restated from published algorithms only: diffrax's stepping loop and jax.random.
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("MTGP_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "tools", "refshim"))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

import jax  # noqa: E402  (the shim)
import jax.numpy as jnp  # noqa: E402

assert getattr(jax, "SHIM", False), "this recorder is for images WITHOUT jax; use tools/record_reference_golden.py"

from MultiTreeGP.environments.control_environments.acrobot import Acrobot  # noqa: E402
from MultiTreeGP.environments.control_environments.harmonic_oscillator import HarmonicOscillator  # noqa: E402
from MultiTreeGP.environments.control_environments.reactor import StirredTankReactor  # noqa: E402
from MultiTreeGP.genetic_programming import GeneticProgramming  # noqa: E402

from bench_support.synthetic import sample_population  # noqa: E402
from multitreegp_b200.node_library import NodeLibrary  # noqa: E402

F = np.float32
OUT = os.path.join(ROOT, "tests", "golden", "refshim")


def A(x):
    return jnp.asarray(np.asarray(x, dtype=F))


# ---- trees ---------------------------------------------------------------------------------------
def tree_cases():
    ops_dyn = [("+", lambda x, y: x + y, 2, 0.5), ("-", lambda x, y: x - y, 2, 0.1), ("*", lambda x, y: x * y, 2, 0.5),
               ("sin", lambda x: jnp.sin(x), 1, 0.1), ("cos", lambda x: jnp.cos(x), 1, 0.1)]  # examples/DynamicPolicy.ipynb
    vars_dyn = [["y1", "y2", "y3", "y4", "a1", "a2", "u"], ["a1", "a2"]]
    ops_arith = [("+", lambda x, y: x + y, 2, 0.5), ("-", lambda x, y: x - y, 2, 0.2), ("*", lambda x, y: x * y, 2, 0.5),
                 ("/", lambda x, y: x / y, 2, 0.2)]  # IEEE-exact operators only: compared bit for bit
    vars_sr = [["x0", "x1"]]
    return [("dynamic", ops_dyn, vars_dyn, [2, 1], 4, 30, 7), ("arith", ops_arith, vars_sr, [2], 5, 40, 2)]


def record_trees():
    out = {}
    rng = np.random.default_rng(20261019)
    for name, ops, variables, layers, depth, max_nodes, n_data in tree_cases():
        gp = GeneticProgramming(2, 20, (lambda *a, **k: 0.0), ops, variables, jnp.array(layers), max_init_depth=depth,
                                max_nodes=max_nodes)
        lib = NodeLibrary(ops, variables, layers)
        out[f"{name}_node_strings"] = np.array(sorted(gp.string_to_node, key=gp.string_to_node.get))
        out[f"{name}_node_ids"] = np.array([gp.string_to_node[s] for s in out[f"{name}_node_strings"]], dtype=np.int32)
        out[f"{name}_slots"] = np.asarray(gp.slots, dtype=np.int32)
        pop = sample_population(rng, 48, lib, max_init_depth=depth, max_nodes=max_nodes)
        pop = np.concatenate([pop, sample_population(rng, 16, lib, max_init_depth=depth, max_nodes=max_nodes, op_growth=0.95)])
        data = rng.normal(size=(6, n_data)).astype(F) * F(1.5)
        data[0] = 0.0  # x / 0, 0 / 0
        vals = np.empty((pop.shape[0], data.shape[0], pop.shape[1]), dtype=F)
        with np.errstate(all="ignore"):
            for p in range(pop.shape[0]):
                for d in range(data.shape[0]):
                    vals[p, d] = np.asarray(gp.vmap_foriloop(A(pop[p]), A(data[d])), dtype=F)
        out[f"{name}_pop"], out[f"{name}_data"], out[f"{name}_values"] = pop, data, vals
    return out


# ---- environments ------------------------------------------------------------------------------
def env_cases(rng):
    acro = [(1.0, 1.0, 1.0, 1.0), (1.1, 0.9, 1.2, 0.8)]
    harm = [(1.0, 0.0), (1.3, 0.4)]
    reac = [(100.0, 239.0, -5.0e4, 5.0e4, 100.0, 300.0, 300.0, 20.0), (120.0, 250.0, -4.8e4, 6.0e4, 90.0, 320.0, 280.0, 25.0)]
    n = 40
    xa = rng.normal(size=(n, 4)).astype(F) * np.array([4.0, 4.0, 6.0, 9.0], dtype=F)  # angles well beyond +-pi
    xh = rng.normal(size=(n, 2)).astype(F) * np.array([3.0, 2.0], dtype=F)
    xr = (np.array([290.0, 360.0, 0.7], dtype=F) + rng.normal(size=(n, 3)).astype(F) * np.array([10.0, 15.0, 0.3], dtype=F)).astype(F)
    return [("acrobot", Acrobot, acro, xa, rng.normal(size=n).astype(F) * F(1.5), 0),
            ("harmonic", HarmonicOscillator, harm, xh, rng.normal(size=n).astype(F) * F(2.0), 1),
            ("reactor", StirredTankReactor, reac, xr, (rng.uniform(-50, 350, size=n)).astype(F), 1)]


def record_env():
    out = {}
    rng = np.random.default_rng(7)
    K = 40
    ts = np.linspace(0.0, 20.0, K).astype(F)
    for name, cls, param_sets, xs, us, n_targets in env_cases(rng):
        D = xs.shape[1]
        drift = np.empty((len(param_sets),) + xs.shape, dtype=F)
        fobs = np.empty_like(drift)
        m = 5
        traj_x = (xs[:m, None, :] + rng.normal(size=(m, K, D)).astype(F) * F(0.5)).astype(F)
        if name == "acrobot":  # some trajectories reach the swing-up threshold, some never do
            traj_x[0, :, 0] = np.linspace(0.0, 3.5, K).astype(F)
            traj_x[0, :, 1] = 0.0
            traj_x[1, :, :2] = F(0.1)
        traj_u = rng.normal(size=(m, K, 1)).astype(F)
        targets = rng.uniform(-3, 3, size=(m, max(n_targets, 1))).astype(F) if name != "reactor" else \
            rng.uniform(400, 500, size=(m, 1)).astype(F)
        fit = np.empty((len(param_sets), m), dtype=F)
        cond_states = np.concatenate([xs[:6], xs[:6] * F(10.0)]).astype(F)
        cond_states[1, 0] = np.nan
        cond_states[2, D - 1] = np.inf
        cond = np.empty(cond_states.shape[0], dtype=F)
        with np.errstate(all="ignore"), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for ip, ps in enumerate(param_sets):
                env = cls(0.0, 0.0)
                env.initialize_parameters(tuple(F(v) for v in ps), A(ts))
                for i in range(xs.shape[0]):
                    drift[ip, i] = np.asarray(env.drift(F(0.3), A(xs[i]), A([us[i]])), dtype=F).reshape(-1)
                    fobs[ip, i] = np.asarray(env.f_obs(None, (F(0.3), A(xs[i])))[1], dtype=F).reshape(-1)
                for j in range(m):
                    tgt = A(targets[j][:n_targets]) if n_targets else A(np.zeros(0))
                    fit[ip, j] = F(env.fitness_function(A(traj_x[j]), A(traj_u[j]), tgt, A(ts)))
            env = cls(0.0, 0.0)
            for i in range(cond_states.shape[0]):
                cond[i] = F(env.cond_fn_nan(F(0.0), A(cond_states[i]), None))
        out.update({f"{name}_params": np.asarray(param_sets, dtype=F), f"{name}_x": xs, f"{name}_u": us,
                    f"{name}_drift": drift, f"{name}_f_obs": fobs, f"{name}_traj_x": traj_x, f"{name}_traj_u": traj_u,
                    f"{name}_targets": targets[:, :max(n_targets, 1)], f"{name}_fitness": fit,
                    f"{name}_cond_states": cond_states, f"{name}_cond": cond})
    out["ts"] = ts
    return out


# ---- evaluate_population: the reference's own wrappers around a fitness function ---------------------
def record_population():
    """GeneticProgramming.evaluate_population (genetic_programming.py:403-433) end to end -- reshape of the islands,
    jit_eval = jit(shard_map(vmap(partial_ff))) (:249-269; jit / shard_map are inert here, vmap is a loop), the split
    into (coefficients, nodes), the parsimony term (:424), best-of-generation (:426-431) -- around a fitness function
    with the signature the evaluators have.  The function is the pointwise regression fitness of SURVEY.md App. G,
    written with the reference's primitives (tree_evaluator = vmap_foriloop)."""
    rng = np.random.default_rng(99)
    ops = [("+", lambda x, y: x + y, 2, 0.5), ("-", lambda x, y: x - y, 2, 0.2), ("*", lambda x, y: x * y, 2, 0.5),
           ("/", lambda x, y: x / y, 2, 0.1), ("sin", lambda x: jnp.sin(x), 1, 0.1)]
    variables, layers, depth, max_nodes = [["x0", "x1"]], [2], 4, 24
    num_pop, pop_size, parsimony = 2, 20, 0.003

    def fitness_function(coefficients, nodes, data, tree_evaluator):
        candidate = jnp.concatenate([nodes, coefficients], axis=-1)  # as SR_evaluator.py:33 / feedforward_evaluate.py:39 do
        X, Y = data
        pred = jax.vmap(tree_evaluator, in_axes=[None, 0])(candidate, X)
        fitness = jnp.mean(jnp.sum(jnp.square(pred - Y), axis=-1))
        fitness = jnp.where(jnp.isnan(fitness) | jnp.isinf(fitness), F(1e5), fitness)
        return jnp.clip(fitness, 0, 1e5)

    gp = GeneticProgramming(3, pop_size, fitness_function, ops, variables, jnp.array(layers), num_populations=num_pop,
                            max_init_depth=depth, max_nodes=max_nodes, size_parsinomy=parsimony)
    lib = NodeLibrary(ops, variables, layers)
    pops = sample_population(rng, num_pop * pop_size, lib, max_init_depth=depth, max_nodes=max_nodes)
    pops = pops.reshape(num_pop, pop_size, *pops.shape[1:])
    X = rng.normal(size=(64, 2)).astype(F)
    Y = np.stack([X[:, 0] * X[:, 1] + F(0.5), np.sin(X[:, 0]) - X[:, 1]], axis=1).astype(F)
    with np.errstate(all="ignore"):
        fitness, same = gp.evaluate_population(A(pops), (A(X), A(Y)))
    return dict(pop_populations=pops, pop_X=X, pop_Y=Y, pop_fitness=np.asarray(fitness, dtype=F),
                pop_parsimony=np.array(parsimony, dtype=F), pop_best_fitness=np.asarray(gp.best_fitnesses[0], dtype=F),
                pop_best_solution=np.asarray(gp.best_solutions[0], dtype=F),
                pop_returned_population_is_input=np.array(np.array_equal(np.asarray(same), pops)))


def main():
    meta = dict(recorder="tools/record_reference_shim_golden.py", reference=REF, numpy=np.__version__,
                note="reference code executed over tools/refshim (numpy float32), noise-free f_obs")
    trees, env = record_trees(), record_env()
    trees.update(record_population())
    np.savez_compressed(os.path.join(OUT, "refshim_trees.npz"), meta=np.array(repr(meta)), **trees)
    np.savez_compressed(os.path.join(OUT, "refshim_env.npz"), meta=np.array(repr(meta)), **env)
    print("recorded", sorted(trees), sorted(env))


if __name__ == "__main__":
    main()
