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


# ---- the evaluators around diffrax.diffeqsolve ----------------------------------------------------------
def record_evaluators():
    """feedforward_evaluate.Evaluator / dynamic_evaluate.Evaluator (the reference's classes, unmodified) with the ONE
    third-party call they make -- diffrax.diffeqsolve -- replaced by an injection of the saved states the oracle's own
    solve returns for the same inputs.  Everything else is the reference executing: the vmap over trajectories and the
    data tuple layout (evaluate_candidate), env.initialize_parameters, the initial state handed to the solver, the
    post-pass over the save points (f_obs scan, readout, dynamic_evaluate.py:98-103 / feedforward_evaluate.py:95-99),
    env.fitness_function, the NaN / inf penalty + mean + clip of __call__ (:48-51), and _drift (the vector field,
    evaluated directly at sampled points).  Observation noise is off (jax.random is third-party)."""
    import diffrax  # the shim
    from MultiTreeGP.evaluators import dynamic_evaluate, feedforward_evaluate
    from multitreegp_b200 import abi
    from oracle import pyoracle
    from tests import util
    pyoracle.build()
    ref_env = {"acrobot": Acrobot, "harmonic": HarmonicOscillator, "reactor": StirredTankReactor}
    ops_ref = [("+", lambda x, y: x + y, 2, 0.5), ("-", lambda x, y: x - y, 2, 0.1), ("*", lambda x, y: x * y, 2, 0.5),
               ("sin", lambda x: jnp.sin(x), 1, 0.1), ("cos", lambda x: jnp.cos(x), 1, 0.1)]
    out = {}
    rng = np.random.default_rng(314)
    for name, env_name, kind in (("static_harmonic", "harmonic", abi.EVAL_STATIC), ("static_acrobot", "acrobot", abi.EVAL_STATIC),
                                 ("dynamic_acrobot", "acrobot", abi.EVAL_DYNAMIC), ("dynamic_reactor", "reactor", abi.EVAL_DYNAMIC)):
        P, B, S = 6, 4, (2 if kind == abi.EVAL_DYNAMIC else 0)
        lib, pop, data = util.control_case(env_name, kind, rng, P=P, B=B, T_end=4.0, dt_save=0.2, dt0=0.05, obs_noise=0.0,
                                           param_mode="Different" if env_name != "acrobot" else "Constant")
        _, variables, layers = util.control_library_args(util.ENVS[env_name](0.05, 0.0), kind, 2)
        gp = GeneticProgramming(2, 20, (lambda *a, **k: 0.0), ops_ref, variables, jnp.array(layers), max_init_depth=4, max_nodes=30)
        assert gp.string_to_node == lib.string_to_node
        env = ref_env[env_name](0.05, 0.0)
        ev = (dynamic_evaluate.Evaluator(env, S, 0.05) if kind == abi.EVAL_DYNAMIC else feedforward_evaluate.Evaluator(env, 0.05))
        orc = pyoracle.rollout_population(pop, lib.table(), data, trajectories=True)
        sol_ys = np.concatenate([orc["xs"], orc["activities"]], axis=-1) if kind == abi.EVAL_DYNAMIC else orc["xs"]
        D = data.x0.shape[1]
        keys = np.zeros((B, 2), np.uint32)
        params = tuple(A(data.params[:, i]) for i in range(data.params.shape[1]))
        ref_data = (A(data.x0), A(data.ts), A(data.targets), jnp.asarray(keys), jnp.asarray(keys), params)
        r_ys = np.empty((P, B, data.K, env.n_obs), F); r_us = np.empty((P, B, data.K, 1), F)
        r_fit = np.empty((P, B), F); r_call = np.empty(P, F); y0s = np.empty((P, B, D + S), F)
        n_pts = 5
        pts = np.empty((P, n_pts, D + S), F); drifts = np.empty((P, n_pts, D + S), F)
        with np.errstate(all="ignore"), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for p in range(P):
                diffrax.INJECT[:] = [sol_ys[p, b] for b in range(B)]
                diffrax.CALLS.clear()
                res = ev.evaluate_candidate(A(pop[p]), ref_data, gp.vmap_foriloop)
                xs_r, ys_r, us_r, fit_r = res[0], res[1], res[2], res[-1]
                assert np.array_equal(np.asarray(xs_r), orc["xs"][p], equal_nan=True)
                r_ys[p], r_us[p], r_fit[p] = np.asarray(ys_r), np.asarray(us_r).reshape(B, data.K, 1), np.asarray(fit_r)
                for b in range(B):
                    c = diffrax.CALLS[b]
                    y0s[p, b] = c["y0"]
                    assert c["t0"] == float(data.ts[0]) and c["t1"] == float(data.ts[-1]) and c["dt0"] == F(0.05)
                    assert c["max_steps"] == 16 ** 4 and c["throw"] is False and c["has_event"]
                diffrax.INJECT[:] = [sol_ys[p, b] for b in range(B)]
                r_call[p] = F(ev(A(pop[p][..., 3:]), A(pop[p][..., :3]), ref_data, gp.vmap_foriloop))
                # the vector field at sampled points of trajectory 0's neighbourhood
                e0 = copy_env(ev.env, tuple(q[0] for q in params), ref_data[1])
                tgt = ref_data[2][0]
                for k in range(n_pts):
                    pt = (np.concatenate([data.x0[0], np.zeros(S, F)]) + rng.normal(size=D + S).astype(F) * F(0.7)).astype(F)
                    args = ((e0, A(pop[p][:S]), A(pop[p][S:]), None, tgt, gp.vmap_foriloop) if kind == abi.EVAL_DYNAMIC
                            else (e0, A(pop[p]), None, tgt, gp.vmap_foriloop))
                    pts[p, k] = pt
                    drifts[p, k] = np.asarray(ev._drift(F(0.25), A(pt), args), dtype=F).reshape(-1)
        out.update({f"ev_{name}_pop": pop, f"ev_{name}_x0": data.x0, f"ev_{name}_ts": data.ts, f"ev_{name}_targets": data.targets,
                    f"ev_{name}_params": data.params, f"ev_{name}_sol_ys": sol_ys, f"ev_{name}_y0": y0s, f"ev_{name}_ys": r_ys,
                    f"ev_{name}_us": r_us, f"ev_{name}_traj_fitness": r_fit, f"ev_{name}_fitness": r_call,
                    f"ev_{name}_points": pts, f"ev_{name}_drift": drifts})
    # --- SR_evaluator.Evaluator: evaluate_time_series around the solve (MSE fitness), __call__, _drift ---
    from MultiTreeGP.evaluators import SR_evaluator
    lib, pop, data = util.sr_case(rng, P=8, B=4, T_end=3.0)
    ops_sr = [("+", lambda x, y: x + y, 2, 0.5), ("-", lambda x, y: x - y, 2, 0.1), ("*", lambda x, y: x * y, 2, 0.5),
              ("/", lambda x, y: x / y, 2, 0.1)]
    gp = GeneticProgramming(2, 20, (lambda *a, **k: 0.0), ops_sr, [["x0", "x1"]], jnp.array([2]), max_init_depth=4, max_nodes=30)
    assert gp.string_to_node == lib.string_to_node
    ev = SR_evaluator.Evaluator(dt0=float(data.dt0))
    orc = pyoracle.rollout_population(pop, lib.table(), data, trajectories=True)
    P, B = pop.shape[0], data.x0.shape[0]
    keys = np.zeros((B, 2), np.uint32)
    ref_data = (A(data.x0), A(data.ts), A(data.ys_true), jnp.asarray(keys))
    r_fit = np.empty((P, B), F); r_call = np.empty(P, F)
    pts = (rng.normal(size=(P, 5, 2)) * 1.5).astype(F); drifts = np.empty_like(pts)
    with np.errstate(all="ignore"), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for p in range(P):
            diffrax.INJECT[:] = [orc["xs"][p, b] for b in range(B)]
            diffrax.CALLS.clear()
            fit_r, pred = ev.evaluate_candidate(A(pop[p]), ref_data, gp.vmap_foriloop)
            assert np.array_equal(np.asarray(pred), orc["xs"][p], equal_nan=True)
            assert all(np.array_equal(c["y0"], data.x0[b]) and c["throw"] is False and c["has_event"] for b, c in enumerate(diffrax.CALLS))
            r_fit[p] = np.asarray(fit_r)
            diffrax.INJECT[:] = [orc["xs"][p, b] for b in range(B)]
            r_call[p] = F(ev(A(pop[p][..., 3:]), A(pop[p][..., :3]), ref_data, gp.vmap_foriloop))
            for k in range(pts.shape[1]):
                drifts[p, k] = np.asarray(ev._drift(F(0.0), A(pts[p, k]), (A(pop[p]), gp.vmap_foriloop)), dtype=F)
    out.update(ev_sr_pop=pop, ev_sr_x0=data.x0, ev_sr_ts=data.ts, ev_sr_ys_true=data.ys_true, ev_sr_dt0=np.array(data.dt0, F),
               ev_sr_sol_ys=orc["xs"], ev_sr_traj_fitness=r_fit, ev_sr_fitness=r_call, ev_sr_points=pts, ev_sr_drift=drifts)
    return out


def copy_env(env, params, ts):
    import copy
    e = copy.copy(env)
    e.initialize_parameters(params, ts)
    return e


def main():
    meta = dict(recorder="tools/record_reference_shim_golden.py", reference=REF, numpy=np.__version__,
                note="reference code executed over tools/refshim (numpy float32), noise-free f_obs")
    trees, env = record_trees(), record_env()
    trees.update(record_population())
    np.savez_compressed(os.path.join(OUT, "refshim_evaluators.npz"), meta=np.array(repr(meta)), **record_evaluators())
    np.savez_compressed(os.path.join(OUT, "refshim_trees.npz"), meta=np.array(repr(meta)), **trees)
    np.savez_compressed(os.path.join(OUT, "refshim_env.npz"), meta=np.array(repr(meta)), **env)
    print("recorded", sorted(trees), sorted(env))


if __name__ == "__main__":
    main()
