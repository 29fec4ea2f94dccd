"""CPU oracle vs fixtures recorded from the reference's OWN code (tools/record_reference_shim_golden.py).

The fixtures hold outputs of /root/reference/MultiTreeGP itself -- GeneticProgramming.__init__ / vmap_foriloop and the
three environment classes -- executed over tools/refshim (numpy float32 stand-in for jax).  They pin everything of the
hot path that is the reference's own code; diffrax's loop and jax.random stay restated from their published algorithms
(tests/test_reference_golden.py takes over on a machine with jax + diffrax).
"""
import ctypes as C
import os

import numpy as np
import pytest

from multitreegp_b200 import abi
from multitreegp_b200.node_library import NodeLibrary
from multitreegp_b200.rollout_data import RolloutData
from oracle import pyoracle

HERE = os.path.dirname(os.path.abspath(__file__))
RTOL = 1e-5  # BASELINE.json north_star: rtol 1e-5 against the reference


def _load(name):
    path = os.path.join(HERE, "golden", "refshim", name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not recorded")
    return np.load(path, allow_pickle=False)


TREE_CASES = {
    "dynamic": ([("+", 2, 0.5), ("-", 2, 0.1), ("*", 2, 0.5), ("sin", 1, 0.1), ("cos", 1, 0.1)],
                [["y1", "y2", "y3", "y4", "a1", "a2", "u"], ["a1", "a2"]], [2, 1]),
    "arith": ([("+", 2, 0.5), ("-", 2, 0.2), ("*", 2, 0.5), ("/", 2, 0.2)], [["x0", "x1"]], [2]),
}


@pytest.mark.parametrize("case", sorted(TREE_CASES))
def test_node_library_equals_reference_constructor(case):
    """NodeLibrary == the id assignment GeneticProgramming.__init__ itself produced (genetic_programming.py:132-199)."""
    g = _load("refshim_trees.npz")
    ops, variables, layers = TREE_CASES[case]
    lib = NodeLibrary(ops, variables, layers)
    ref = dict(zip((str(s) for s in g[f"{case}_node_strings"]), (int(i) for i in g[f"{case}_node_ids"])))
    assert lib.string_to_node == ref
    assert np.array_equal(lib.slots, g[f"{case}_slots"])


@pytest.mark.parametrize("case", sorted(TREE_CASES))
def test_tree_values_equal_reference_interpreter(case):
    """oracle_tree_eval == GeneticProgramming.vmap_foriloop (body_fun, :356-401) on the recorded trees."""
    g = _load("refshim_trees.npz")
    ops, variables, layers = TREE_CASES[case]
    table = NodeLibrary(ops, variables, layers).table()
    pop, data, ref = g[f"{case}_pop"], g[f"{case}_data"], g[f"{case}_values"]
    got = np.empty_like(ref)
    for p in range(pop.shape[0]):
        for d in range(data.shape[0]):
            for t in range(pop.shape[1]):
                got[p, d, t] = pyoracle.tree_eval(pop[p, t], table, data[d])
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    if case == "arith":  # + - * / are IEEE-exact on both sides: bit for bit (NaN sign / payload aside)
        m = ~np.isnan(ref)
        assert np.array_equal(got[m].view(np.uint32), ref[m].view(np.uint32))
    else:  # numpy's float32 sin / cos vs the contract's: a few ulp
        np.testing.assert_allclose(got, ref, rtol=RTOL, atol=1e-6, equal_nan=True)
    assert np.isfinite(ref).mean() > 0.5


ENVS = {"acrobot": (abi.ENV_ACROBOT, 0), "harmonic": (abi.ENV_HARMONIC, 1), "reactor": (abi.ENV_REACTOR, 1)}


def _desc(env_id, params, D, ts, targets):
    rd = RolloutData(kind=abi.EVAL_STATIC, env_id=env_id, x0=np.zeros((1, D), np.float32), ts=ts, dt0=0.1,
                     max_fitness=1e4, targets=targets, params=np.asarray(params, np.float32).reshape(1, -1))
    return rd.to_desc_host()


@pytest.mark.parametrize("env", sorted(ENVS))
def test_environment_functions_equal_reference_classes(env):
    """drift / f_obs / cond_fn_nan / fitness_function of the oracle == the reference's environment classes."""
    g = _load("refshim_env.npz")
    env_id, n_targets = ENVS[env]
    L = pyoracle.lib()
    L.oracle_env_cond.restype = C.c_float
    L.oracle_env_fitness.restype = C.c_float
    ts = g["ts"]
    xs, us = g[f"{env}_x"], g[f"{env}_u"]
    D = xs.shape[1]
    for ip, ps in enumerate(g[f"{env}_params"]):
        desc, keep = _desc(env_id, ps, D, ts, np.zeros((1, n_targets), np.float32))
        drift, fobs = np.empty_like(xs), np.empty_like(xs)
        for i in range(xs.shape[0]):
            x = np.ascontiguousarray(xs[i])
            L.oracle_env_drift(C.byref(desc), 0, C.c_void_p(x.ctypes.data), C.c_float(us[i]), C.c_void_p(drift[i].ctypes.data))
            L.oracle_env_f_obs(C.byref(desc), 0, C.c_float(0.3), C.c_void_p(x.ctypes.data), C.c_void_p(fobs[i].ctypes.data))
        scale = np.abs(g[f"{env}_drift"][ip]).max(axis=0)  # a component that cancels to ~0 is judged on its column's scale
        np.testing.assert_allclose(drift, g[f"{env}_drift"][ip], rtol=RTOL, atol=RTOL * scale.max())
        np.testing.assert_allclose(fobs, g[f"{env}_f_obs"][ip], rtol=RTOL, atol=1e-5)
        fit = np.empty(g[f"{env}_fitness"].shape[1], np.float32)
        for j in range(fit.size):
            tx = np.ascontiguousarray(g[f"{env}_traj_x"][j])
            tu = np.ascontiguousarray(g[f"{env}_traj_u"][j].reshape(-1))
            tg = np.ascontiguousarray(g[f"{env}_targets"][j])
            fit[j] = L.oracle_env_fitness(C.byref(desc), 0, C.c_void_p(tx.ctypes.data), C.c_void_p(tu.ctypes.data),
                                          C.c_void_p(tg.ctypes.data), C.c_void_p(ts.ctypes.data), C.c_int(ts.size))
        np.testing.assert_allclose(fit, g[f"{env}_fitness"][ip], rtol=RTOL)
        del keep
    cs = g[f"{env}_cond_states"]
    cond = np.array([L.oracle_env_cond(C.c_int(env_id), C.c_void_p(np.ascontiguousarray(c).ctypes.data), C.c_int(D)) for c in cs],
                    np.float32)
    assert np.array_equal(cond, g[f"{env}_cond"])
    assert (cond < 0).any() and (cond > 0).any()


POP_OPS = [("+", 2, 0.5), ("-", 2, 0.2), ("*", 2, 0.5), ("/", 2, 0.1), ("sin", 1, 0.1)]


def test_evaluate_population_equals_reference_wrappers():
    """oracle pointwise fitness + size_parsimony * live rows == GeneticProgramming.evaluate_population itself
    (reshape, jit_eval = jit(shard_map(vmap(partial_ff))), coefficient / node split, parsimony :424, best :426-431)."""
    g = _load("refshim_trees.npz")
    lib = NodeLibrary(POP_OPS, [["x0", "x1"]], [2])
    pops = g["pop_populations"]
    flat = pops.reshape(-1, *pops.shape[2:])
    base = pyoracle.pointwise_population(flat, lib.table(), g["pop_X"], g["pop_Y"], max_fitness=1e5)
    fit = base + np.float32(g["pop_parsimony"]) * pyoracle.live_count(flat).astype(np.float32)
    ref = g["pop_fitness"].reshape(-1)
    np.testing.assert_allclose(fit, ref, rtol=RTOL)
    assert bool(g["pop_returned_population_is_input"])
    best = int(np.argmin(ref))
    assert np.array_equal(g["pop_best_solution"], flat[best]) and g["pop_best_fitness"] == ref[best]
    order = np.sort(ref)
    if order[1] - order[0] > RTOL * abs(order[0]):  # the winner is decided beyond the tolerance
        assert int(np.argmin(fit)) == best
    assert len(np.unique(ref)) > len(ref) // 2


@pytest.mark.gpu
def test_cuda_evaluate_population_equals_reference_wrappers():
    """The product path (pointwise evaluator -> C ABI -> k_pointwise with the fused parsimony term, EvaluationDriver)
    against the same reference-recorded fitness matrix and best-of-generation."""
    from multitreegp_b200.evaluators import pointwise_evaluate
    from multitreegp_b200.strategy_eval import EvaluationDriver
    g = _load("refshim_trees.npz")
    ev = pointwise_evaluate.Evaluator().bind(POP_OPS, [["x0", "x1"]], [2])
    pops = g["pop_populations"]
    drv = EvaluationDriver(3, pops.shape[1], ev, num_populations=pops.shape[0], size_parsinomy=float(g["pop_parsimony"]))
    fit, same = drv.evaluate_population(pops, (g["pop_X"], g["pop_Y"]))
    assert same is pops
    np.testing.assert_allclose(np.asarray(fit), g["pop_fitness"], rtol=RTOL)
    bf, bs = drv.get_statistics(0)
    ref = g["pop_fitness"].reshape(-1)
    order = np.sort(ref)
    if order[1] - order[0] > RTOL * abs(order[0]):
        assert np.array_equal(np.asarray(bs), g["pop_best_solution"])
    np.testing.assert_allclose(bf, g["pop_best_fitness"], rtol=RTOL)


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(TREE_CASES))
def test_cuda_tree_values_equal_reference_interpreter(case):
    """k_encode -> k_lower -> k_pointwise predictions == GeneticProgramming.vmap_foriloop on the recorded trees."""
    from multitreegp_b200.evaluators import pointwise_evaluate
    g = _load("refshim_trees.npz")
    ops, variables, layers = TREE_CASES[case]
    n_trees = sum(layers)
    ev = pointwise_evaluate.Evaluator().bind(ops, [sorted({v for vl in variables for v in vl}, key=[v for vl in variables for v in vl].index)], [n_trees])
    pop, data, ref = g[f"{case}_pop"], g[f"{case}_data"], g[f"{case}_values"]
    Y = np.zeros((data.shape[0], n_trees), np.float32)
    for p in range(pop.shape[0]):
        _, pred = ev.evaluate_candidate(pop[p], (data, Y))
        if case == "arith":
            m = ~np.isnan(ref[p])
            assert np.array_equal(np.isnan(pred), np.isnan(ref[p]))
            assert np.array_equal(pred[m].view(np.uint32), ref[p][m].view(np.uint32))
        else:
            np.testing.assert_allclose(pred, ref[p], rtol=RTOL, atol=1e-6, equal_nan=True)


EV_CASES = {"static_harmonic": ("harmonic", abi.EVAL_STATIC), "static_acrobot": ("acrobot", abi.EVAL_STATIC),
            "dynamic_acrobot": ("acrobot", abi.EVAL_DYNAMIC), "dynamic_reactor": ("reactor", abi.EVAL_DYNAMIC)}


def _ev_case(g, name):
    from tests import util
    env_name, kind = EV_CASES[name]
    env = util.ENVS[env_name](0.05, 0.0)
    S = 2 if kind == abi.EVAL_DYNAMIC else 0
    lib = util.control_library(env, kind, 2)
    data = RolloutData(kind=kind, env_id=env.env_id, x0=g[f"ev_{name}_x0"], ts=g[f"ev_{name}_ts"], dt0=0.05, max_fitness=1e4,
                       targets=g[f"ev_{name}_targets"], obs_keys=np.zeros((g[f"ev_{name}_x0"].shape[0], 2), np.uint32),
                       params=g[f"ev_{name}_params"], n_obs=env.n_obs, state_size=S, obs_noise=0.0)
    return lib, g[f"ev_{name}_pop"], data, S


@pytest.mark.parametrize("name", sorted(EV_CASES))
def test_evaluators_around_the_solve_equal_reference_classes(name):
    """The reference's feedforward / dynamic Evaluator classes, run unmodified with diffrax.diffeqsolve replaced by an
    injection of the oracle's saved states: what they hand to the solver, their post-pass over the save points (f_obs,
    readout), env.fitness_function, the penalty + mean + clip of __call__ and _drift -- against the oracle's."""
    g = _load("refshim_evaluators.npz")
    lib, pop, data, S = _ev_case(g, name)
    orc = pyoracle.rollout_population(pop, lib.table(), data, trajectories=True)
    sol = np.concatenate([orc["xs"], orc["activities"]], axis=-1) if S else orc["xs"]
    # the saved states that were injected are the oracle's solve of today (regression pin of the restated solve)
    assert np.array_equal(sol.view(np.uint32), g[f"ev_{name}_sol_ys"].view(np.uint32))
    # what the reference handed to diffeqsolve as y0: x0 (static) / concatenate([x0, zeros(state_size)]) (dynamic_evaluate.py:89)
    want_y0 = np.concatenate([data.x0, np.zeros((data.B, S), np.float32)], axis=1)
    assert np.array_equal(g[f"ev_{name}_y0"], np.broadcast_to(want_y0, g[f"ev_{name}_y0"].shape))
    # post-pass at the save points; rows after an event are the inf pre-fill of SaveAt on both sides
    fin = np.isfinite(g[f"ev_{name}_ys"])
    assert np.array_equal(fin, np.isfinite(orc["ys"]))
    np.testing.assert_allclose(orc["ys"][fin], g[f"ev_{name}_ys"][fin], rtol=RTOL, atol=1e-5)
    fu = np.isfinite(g[f"ev_{name}_us"])
    assert np.array_equal(fu, np.isfinite(orc["us"]))
    np.testing.assert_allclose(orc["us"][fu], g[f"ev_{name}_us"][fu], rtol=RTOL, atol=1e-5)
    ft = np.isfinite(g[f"ev_{name}_traj_fitness"])
    assert np.array_equal(ft, np.isfinite(orc["traj_fitness"]))
    np.testing.assert_allclose(orc["traj_fitness"][ft], g[f"ev_{name}_traj_fitness"][ft], rtol=RTOL)
    np.testing.assert_allclose(orc["fitness"], g[f"ev_{name}_fitness"], rtol=RTOL)
    # the vector field at the sampled points
    L = pyoracle.lib()
    L.oracle_vector_field.restype = C.c_int
    desc, keep = data.to_desc_host()
    pts, ref = g[f"ev_{name}_points"], g[f"ev_{name}_drift"]
    got = np.empty_like(ref)
    table = lib.table()
    for p in range(pop.shape[0]):
        cand = np.ascontiguousarray(pop[p])
        for k in range(pts.shape[1]):
            y = np.ascontiguousarray(pts[p, k])
            assert L.oracle_vector_field(C.c_void_p(cand.ctypes.data), cand.shape[0], cand.shape[1], C.byref(table), C.byref(desc),
                                         0, C.c_float(0.25), C.c_void_p(y.ctypes.data), C.c_void_p(got[p, k].ctypes.data)) == 0
    del keep
    assert np.array_equal(np.isfinite(got), np.isfinite(ref))
    m = np.isfinite(ref)
    scale = np.abs(ref[m]).max()
    np.testing.assert_allclose(got[m], ref[m], rtol=RTOL, atol=1e-6 * scale)


def test_sr_evaluator_around_the_solve_equals_reference_class():
    """SR_evaluator.Evaluator (reference class, unmodified, injected solve): the MSE fitness of evaluate_time_series
    (:24,81-82), the penalty + mean + clip of __call__ (:42-45) and _drift (:85-88) -- against the oracle's."""
    g = _load("refshim_evaluators.npz")
    from tests import util
    lib = NodeLibrary(util.OPS_SR, [["x0", "x1"]], [2])
    pop = g["ev_sr_pop"]
    data = RolloutData(kind=abi.EVAL_SR, env_id=abi.ENV_NONE, x0=g["ev_sr_x0"], ts=g["ev_sr_ts"], dt0=float(g["ev_sr_dt0"]),
                       max_fitness=1e5, ys_true=g["ev_sr_ys_true"])
    orc = pyoracle.rollout_population(pop, lib.table(), data, trajectories=True)
    assert np.array_equal(orc["xs"].view(np.uint32), g["ev_sr_sol_ys"].view(np.uint32))
    ft = np.isfinite(g["ev_sr_traj_fitness"])
    assert np.array_equal(ft, np.isfinite(orc["traj_fitness"])) and ft.any()
    np.testing.assert_allclose(orc["traj_fitness"][ft], g["ev_sr_traj_fitness"][ft], rtol=RTOL)
    np.testing.assert_allclose(orc["fitness"], g["ev_sr_fitness"], rtol=RTOL)
    L = pyoracle.lib()
    L.oracle_vector_field.restype = C.c_int
    desc, keep = data.to_desc_host()
    table = lib.table()
    pts, ref = g["ev_sr_points"], g["ev_sr_drift"]
    got = np.empty_like(ref)
    for p in range(pop.shape[0]):
        cand = np.ascontiguousarray(pop[p])
        for k in range(pts.shape[1]):
            y = np.ascontiguousarray(pts[p, k])
            assert L.oracle_vector_field(C.c_void_p(cand.ctypes.data), cand.shape[0], cand.shape[1], C.byref(table), C.byref(desc),
                                         0, C.c_float(0.0), C.c_void_p(y.ctypes.data), C.c_void_p(got[p, k].ctypes.data)) == 0
    del keep
    m = ~np.isnan(ref)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    assert np.array_equal(got[m].view(np.uint32), ref[m].view(np.uint32))  # + - * / : bit for bit
