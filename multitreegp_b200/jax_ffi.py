"""jax.ffi binding of the CUDA evaluators (the host language of the reference is Python/JAX).

jax is not installable in this image (nor on the GPU boxes of this build), so the jax side of this module
(`register`, `JaxEvaluator.__call__`) can only run on a machine that has jax >= 0.4.38.  What CAN be checked here
is: the C++ handlers compile and run against a stand-in for xla/ffi/api/ffi.h with buffers shaped as
`vmap_method="expand_dims"` delivers them (tests/test_ffi_adapter.py, GPU), and the attribute set this module
sends (`rollout_attributes`) is exactly the handler's (same test, same names and order).

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
    """jax.ffi.register_ffi_target for the two handlers of libmtgp_b200_ffi.so (platform "CUDA")."""
    global _registered
    if _registered:
        return
    import jax  # noqa: F401  (raises ImportError where jax is unavailable -- there is no fallback)
    if not os.path.exists(_FFI_LIB):
        raise RuntimeError(f"{_FFI_LIB} is missing: run `python -m multitreegp_b200.build` on a machine with jax")
    lib = ctypes.CDLL(_FFI_LIB)
    jax.ffi.register_ffi_target("mtgp_rollout", jax.ffi.pycapsule(lib.mtgp_rollout_ffi), platform="CUDA")
    jax.ffi.register_ffi_target("mtgp_pointwise", jax.ffi.pycapsule(lib.mtgp_pointwise_ffi), platform="CUDA")
    _registered = True


def rollout_attributes(ev, threefry_partitionable: bool) -> dict:
    """The attribute set of the "mtgp_rollout" handler (csrc/mtgp_xla_ffi.cc) for a bound evaluator: node table,
    evaluator kind, environment, solver and step-size controller incl. the PID gains and limits."""
    from . import abi
    lib, ss = ev.library, ev.solver_settings()
    if ev.kind == abi.EVAL_SR:
        env_id, n_obs, state_size, obs_noise = abi.ENV_NONE, 0, 0, 0.0
    else:
        env_id, n_obs, obs_noise = ev.env.env_id, ev.env.n_obs, float(ev.env.obs_noise)
        state_size = getattr(ev, "state_size", 0)
    pid = ss.get("pid", (0.0, 0.0, 0.0))
    lim = ss.get("pid_limits", (0.0, 0.0, 0.0))
    f32, i32 = np.float32, np.int32
    return dict(opcodes=np.asarray(lib.opcodes, np.int32), args=np.asarray(lib.args, np.int32),
                kind=i32(ev.kind), env_id=i32(env_id), n_obs=i32(n_obs), state_size=i32(state_size),
                max_steps=i32(ev.max_steps), dt0=f32(ev.dt0), max_fitness=f32(ev.max_fitness), obs_noise=f32(obs_noise),
                threefry_partitionable=i32(bool(threefry_partitionable)),
                solver=i32(ss.get("solver", 0)), controller=i32(ss.get("controller", 0)),
                rtol=f32(ss.get("rtol", 0.0)), atol=f32(ss.get("atol", 0.0)), dtmin=f32(ss.get("dtmin", 0.0)),
                pcoeff=f32(pid[0]), icoeff=f32(pid[1]), dcoeff=f32(pid[2]),
                pid_safety=f32(lim[0]), pid_factormin=f32(lim[1]), pid_factormax=f32(lim[2]))


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
        from .evaluators import pointwise_evaluate
        ev = self.ev
        if isinstance(ev, pointwise_evaluate.Evaluator):
            X, Y = data
            call = jax.ffi.ffi_call("mtgp_pointwise", jax.ShapeDtypeStruct((), jnp.float32), vmap_method="expand_dims")
            return call(coefficients, nodes, X, Y.reshape(X.shape[0], -1), opcodes=np.asarray(ev.library.opcodes, np.int32),
                        args=np.asarray(ev.library.args, np.int32), max_fitness=np.float32(ev.max_fitness))
        if ev.kind == abi.EVAL_SR:
            x0, ts, ys, _ = data
            # SR has no observation keys, targets or parameters: zero-size operands (legal for the handler)
            targets, keys, params = jnp.zeros((x0.shape[0], 0)), jnp.zeros((x0.shape[0], 0), jnp.uint32), jnp.zeros((x0.shape[0], 0))
        else:
            x0, ts, targets, _, keys, params = data
            params = jnp.stack(jax.tree_util.tree_leaves(params), axis=1)   # pytree of [B] -> [B, n_params]
            keys = jax.random.key_data(keys) if jnp.issubdtype(keys.dtype, jax.dtypes.prng_key) else keys
            ys = jnp.zeros((0,))
        attrs = rollout_attributes(ev, bool(jax.config.jax_threefry_partitionable))
        if ev.kind == abi.EVAL_SR:
            attrs["n_obs"] = np.int32(x0.shape[1])
        call = jax.ffi.ffi_call("mtgp_rollout", jax.ShapeDtypeStruct((), jnp.float32), vmap_method="expand_dims")
        return call(coefficients, nodes, x0, ts, targets, keys.astype(jnp.uint32), params, ys, **attrs)
