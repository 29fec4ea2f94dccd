"""jax.ffi binding of the CUDA evaluators (the host language of the reference is Python/JAX).

UNTESTED IN THIS IMAGE: jax is not installed here (nor on the GPU boxes of this build), so this
module can only be exercised on a machine that has jax >= 0.4.38.  Everything below the FFI call --
argument validation, encode, lower, fused rollout -- is the C ABI that the GPU parity tests drive
through ctypes, so the adapter carries no logic of its own.

Usage with the UNMODIFIED reference strategy::

    from multitreegp_b200.jax_ffi import JaxEvaluator
    fitness_function = JaxEvaluator(dynamic_evaluate.Evaluator(env, state_size, 0.05)
                                    .bind(operator_list, variable_list, layer_sizes))
    strategy = GeneticProgramming(num_generations, population_size, fitness_function,
                                  operator_list, variable_list, layer_sizes, device_type="gpu")
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_FFI_LIB = os.path.join(_HERE, "libmtgp_b200_ffi.so")
_registered = False


def register() -> None:
    """jax.ffi.register_ffi_target("mtgp_rollout", pycapsule(mtgp_rollout_ffi), platform="CUDA")"""
    global _registered
    if _registered:
        return
    import jax  # noqa: F401  (raises ImportError where jax is unavailable -- there is no fallback)
    if not os.path.exists(_FFI_LIB):
        raise RuntimeError(f"{_FFI_LIB} is missing: run `python -m multitreegp_b200.build` on a machine with jax")
    lib = ctypes.CDLL(_FFI_LIB)
    jax.ffi.register_ffi_target("mtgp_rollout", jax.ffi.pycapsule(lib.mtgp_rollout_ffi), platform="CUDA")
    _registered = True


class JaxEvaluator:
    """Wraps a bound multitreegp_b200 evaluator into the reference's callable protocol for use under
    jax.vmap / shard_map / jit: ``__call__(coefficients, nodes, data, tree_evaluator) -> f32[]``."""

    def __init__(self, evaluator) -> None:
        register()
        self.ev = evaluator
        self.max_fitness = evaluator.max_fitness

    def __call__(self, coefficients, nodes, data, tree_evaluator=None):
        import jax
        import jax.numpy as jnp
        from . import abi
        ev, lib = self.ev, self.ev.library
        ss = ev._solver_settings   # Euler/Dopri5, ConstantStepSize/PIDController (diffrax objects are accepted)
        if ev.kind == abi.EVAL_SR:
            x0, ts, ys, _ = data
            targets, keys, params = jnp.zeros((x0.shape[0], 0)), jnp.zeros((0, 2), jnp.uint32), jnp.zeros((0, 0))
            env_id, n_obs, state_size, obs_noise = abi.ENV_NONE, x0.shape[1], 0, 0.0
        else:
            x0, ts, targets, _, keys, params = data
            params = jnp.stack(jax.tree_util.tree_leaves(params), axis=1)   # pytree of [B] -> [B, n_params]
            keys = jax.random.key_data(keys) if jnp.issubdtype(keys.dtype, jax.dtypes.prng_key) else keys
            ys = jnp.zeros((0,))
            env_id, n_obs, obs_noise = ev.env.env_id, ev.env.n_obs, float(ev.env.obs_noise)
            state_size = getattr(ev, "state_size", 0)
        call = jax.ffi.ffi_call("mtgp_rollout", jax.ShapeDtypeStruct((), jnp.float32), vmap_method="expand_dims")
        return call(coefficients, nodes, x0, ts, targets, keys.astype(jnp.uint32), params, ys,
                    opcodes=np.asarray(lib.opcodes, np.int32), args=np.asarray(lib.args, np.int32),
                    kind=np.int32(ev.kind), env_id=np.int32(env_id), n_obs=np.int32(n_obs),
                    state_size=np.int32(state_size), max_steps=np.int32(ev.max_steps), dt0=np.float32(ev.dt0),
                    max_fitness=np.float32(ev.max_fitness), obs_noise=np.float32(obs_noise),
                    threefry_partitionable=np.int32(bool(jax.config.jax_threefry_partitionable)),
                    solver=np.int32(ss.get("solver", 0)), controller=np.int32(ss.get("controller", 0)),
                    rtol=np.float32(ss.get("rtol", 0.0)), atol=np.float32(ss.get("atol", 0.0)),
                    dtmin=np.float32(ss.get("dtmin", 0.0)))
