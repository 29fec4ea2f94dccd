"""Builders shared by the tests: node libraries, synthetic populations and data of the reference's shapes."""
from __future__ import annotations

import numpy as np

from multitreegp_b200 import abi
from multitreegp_b200.environments.control_environments import Acrobot, HarmonicOscillator, StirredTankReactor
from multitreegp_b200.node_library import NodeLibrary
from multitreegp_b200.rollout_data import RolloutData
from bench_support.synthetic import random_keys, sample_population, van_der_pol_data

OPS_CONTROL = [("+", 2, 0.5), ("-", 2, 0.1), ("*", 2, 0.5), ("sin", 1, 0.1), ("cos", 1, 0.1)]  # StaticPolicy.ipynb:91-95
OPS_SR = [("+", 2, 0.5), ("-", 2, 0.1), ("*", 2, 0.5), ("/", 2, 0.1)]                          # SymbolicRegression.ipynb:138-141
OPS_ALL = OPS_SR + [("sin", 1, 0.1), ("cos", 1, 0.1), ("exp", 1, 0.05), ("log", 1, 0.05), ("sqrt", 1, 0.05),
                    ("tanh", 1, 0.05), ("neg", 1, 0.05), ("abs", 1, 0.05), ("square", 1, 0.05), ("max", 2, 0.05),
                    ("min", 2, 0.05)]

ENVS = {"acrobot": Acrobot, "harmonic": HarmonicOscillator, "reactor": StirredTankReactor}


def control_library_args(env, kind: int, state_size: int = 2, ops=OPS_CONTROL):
    """(operator_list, variable_list, layer_sizes) as the reference's GeneticProgramming constructor takes them."""
    obs = [f"y{i + 1}" for i in range(env.n_obs)]
    tar = [f"tar{i + 1}" for i in range(env.n_targets)]
    if kind == abi.EVAL_STATIC:
        return ops, [obs + tar], [env.n_control]
    lat = [f"a{i + 1}" for i in range(state_size)]
    # DynamicPolicy.ipynb:100-102: state trees see [y, a, u(, target)], the readout sees [a(, target)]
    return ops, [obs + lat + ["u"] + tar, lat + tar], [state_size, env.n_control]


def control_library(env, kind: int, state_size: int = 2, ops=OPS_CONTROL) -> NodeLibrary:
    return NodeLibrary(*control_library_args(env, kind, state_size, ops))


def control_data(env, kind: int, rng, B: int, T_end: float, dt_save: float, dt0: float, state_size: int = 0,
                 param_mode: str = "Constant", max_steps: int = 16 ** 4, partitionable: bool = True) -> RolloutData:
    x0, targets = env.sample_init_states(B, rng)
    ts = np.arange(0, T_end, dt_save, dtype=np.float32)
    params = env.sample_params(B, param_mode, ts, rng)
    params = np.stack([np.asarray(p, dtype=np.float32) for p in params], axis=1)
    return RolloutData(kind=kind, env_id=env.env_id, x0=x0, ts=ts, dt0=dt0, max_fitness=1e4, targets=targets,
                       obs_keys=random_keys(rng, B), params=params, n_obs=env.n_obs, state_size=state_size,
                       max_steps=max_steps, obs_noise=float(env.obs_noise), threefry_partitionable=partitionable)


def sr_case(rng, P: int, B: int = 8, T_end: float = 4.0, dt0: float = 0.01, depth: int = 4, N: int = 30):
    lib = NodeLibrary(OPS_SR, [["x0", "x1"]], [2])
    pop = sample_population(rng, P, lib, depth, N)
    x0, ts, ys, _ = van_der_pol_data(rng, B, T_end, 0.2)
    data = RolloutData(kind=abi.EVAL_SR, env_id=abi.ENV_NONE, x0=x0, ts=ts, dt0=dt0, max_fitness=1e5, ys_true=ys)
    return lib, pop, data


def control_case(env_name: str, kind: int, rng, P: int, B: int = 8, T_end: float = 5.0, dt_save: float = 0.2,
                 dt0: float = 0.05, obs_noise: float = 0.0, state_size: int = 2, depth: int = 4, N: int = 30,
                 param_mode: str = "Constant", **kw):
    env = ENVS[env_name](0.05, obs_noise)
    S = state_size if kind == abi.EVAL_DYNAMIC else 0
    lib = control_library(env, kind, state_size)
    pop = sample_population(rng, P, lib, depth, N)
    data = control_data(env, kind, rng, B, T_end, dt_save, dt0, S, param_mode, **kw)
    return lib, pop, data


DOPRI_PID = dict(solver=abi.SOLVER_DOPRI5, controller=abi.CONTROLLER_PID, rtol=1e-4, atol=1e-4, dtmin=0.001)  # StaticPolicy.ipynb:102


def with_solver(data: RolloutData, **kw) -> RolloutData:
    for k, v in kw.items():
        setattr(data, k, v)
    return data


def same_or_both_nan(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
