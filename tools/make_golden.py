"""Generates tests/golden/*.npz -- frozen input/output vectors of the hot path.

The reference (JAX/diffrax) cannot be imported in this image, so these vectors are produced by the
CPU oracle (oracle/mtgp_oracle.c, contract math) and are REGRESSION PINS of this repository, not
outputs of the reference: they freeze today's semantics so that neither the oracle nor the kernels
can drift silently.  If a machine with JAX becomes available, regenerate them there with the real
reference and record the jax/diffrax versions.

    python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from multitreegp_b200 import abi  # noqa: E402
from oracle import pyoracle as po  # noqa: E402
from tests import util  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def save_case(name, lib, pop, data, out):
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"), pop=pop, opcodes=lib.opcodes, args=lib.args,
        kind=data.kind, env_id=data.env_id, x0=data.x0, ts=data.ts, dt0=np.float32(data.dt0),
        max_fitness=np.float32(data.max_fitness), targets=data.targets,
        obs_keys=data.obs_keys if data.obs_keys is not None else np.zeros((0, 2), np.uint32),
        params=data.params if data.params is not None else np.zeros((0, 0), np.float32),
        ys_true=data.ys_true if data.ys_true is not None else np.zeros((0,), np.float32),
        n_obs=data.n_obs, state_size=data.state_size, max_steps=data.max_steps, obs_noise=np.float32(data.obs_noise),
        partitionable=int(data.threefry_partitionable), solver=data.solver, controller=data.controller,
        rtol=np.float32(data.rtol), atol=np.float32(data.atol), dtmin=np.float32(data.dtmin), fitness=out["fitness"], traj_fitness=out["traj_fitness"],
        steps=out["steps"])


def golden_cases():
    """The case list, in a fixed order on one RNG stream: (name, lib, pop, data, env_name, library_args).
    tools/record_reference_golden.py replays the same list through the real reference."""
    rng = np.random.default_rng(20261018)
    lib, pop, data = util.sr_case(rng, P=48, B=6, T_end=3.0, dt0=0.01)
    yield "sr_vdp", lib, pop, data, None, (util.OPS_SR, [["x0", "x1"]], [2])
    for env, grid in (("harmonic", dict(T_end=5.0, dt_save=0.2, dt0=0.05)),
                      ("acrobot", dict(T_end=5.0, dt_save=0.2, dt0=0.05)),
                      ("reactor", dict(T_end=1.0, dt_save=0.04, dt0=0.002))):
        for kind, kname in ((abi.EVAL_STATIC, "static"), (abi.EVAL_DYNAMIC, "dynamic")):
            lib, pop, data = util.control_case(env, kind, rng, P=48, B=6, obs_noise=0.1,
                                               param_mode="Constant" if env == "acrobot" else "Different", **grid)
            yield f"{kname}_{env}", lib, pop, data, env, util.control_library_args(util.ENVS[env](0.05, 0.1), kind, 2)
    # the notebooks' solver settings: Dopri5 + PIDController(1e-4, 1e-4, dtmin=0.001), max_steps 1000
    lib, pop, data = util.control_case("acrobot", abi.EVAL_DYNAMIC, rng, P=32, B=6, T_end=6.0, dt_save=0.2, dt0=0.05,
                                       obs_noise=0.1, max_steps=1000)
    util.with_solver(data, **util.DOPRI_PID)
    yield ("dynamic_acrobot_dopri5_pid", lib, pop, data, "acrobot",
           util.control_library_args(util.ENVS["acrobot"](0.05, 0.1), abi.EVAL_DYNAMIC, 2))
    # the rng continues into the jax.random vectors of main()
    yield "_rng", rng, None, None, None, None


def main():
    os.makedirs(OUT, exist_ok=True)
    po.build()
    po.set_math(po.MATH_CONTRACT)
    rng = None
    for name, lib, pop, data, _env, _args in golden_cases():
        if name == "_rng":
            rng = lib
            continue
        save_case(name, lib, pop, data, po.rollout_population(pop, lib.table(), data))
    # jax.random vectors (threefry known answers are asserted literally in the tests)
    keys = util.random_keys(rng, 16)
    ts = np.float32(rng.uniform(0, 50, 16))
    normals = np.stack([np.stack([po.normal(po.fold_in(k, int(np.float32(t).view(np.uint32))), n, part)
                                  if True else None for k, t in zip(keys, ts)])
                        for n, part in ((4, True),)])
    np.savez_compressed(os.path.join(OUT, "obs_noise_normal4.npz"), keys=keys, ts=ts, normals=normals[0])
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
