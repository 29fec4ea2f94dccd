"""ctypes front-end of the CPU oracle (oracle/mtgp_oracle.c).

TEST INFRASTRUCTURE ONLY -- importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from multitreegp_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> None:
    """Compiles the oracle with gcc (seconds).  Idempotent."""
    so = os.path.join(_HERE, "libmtgp_oracle.so")
    src = os.path.join(_HERE, "mtgp_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "mtgp_math.h")
    stale = (not os.path.exists(so)) or any(
        os.path.getmtime(f) > os.path.getmtime(so) for f in (src, hdr) if os.path.exists(f))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))


def _cpu_has_fma() -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    flags = line.split()
                    return "fma" in flags and "avx2" in flags
    except OSError:
        pass
    return False


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        name = "libmtgp_oracle_fma.so" if _cpu_has_fma() else "libmtgp_oracle.so"
        path = os.path.join(_HERE, name)
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.oracle_tree_eval.restype = C.c_float
        L.oracle_schedule.restype = C.c_int
        L.oracle_rollout_population.restype = C.c_int
        L.oracle_pointwise_population.restype = C.c_int
        L.oracle_num_threads.restype = C.c_int
        L.oracle_get_math.restype = C.c_int
        _LIB = L
    return _LIB


MATH_CONTRACT, MATH_LIBM = 0, 1
_FN = {"sin": 0, "cos": 1, "exp": 2, "log": 3, "tanh": 4, "log1p": 5, "erfinv": 6, "remainder": 7}


def set_math(mode: int) -> None:
    lib().oracle_set_math(C.c_int(mode))


def math_apply(name: str, x, y=None) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.ascontiguousarray(x if y is None else y, dtype=np.float32)
    out = np.empty_like(x)
    lib().oracle_math_apply(C.c_int(_FN[name]), C.c_void_p(x.ctypes.data), C.c_void_p(y.ctypes.data),
                            C.c_void_p(out.ctypes.data), C.c_int(x.size))
    return out


def threefry2x32(k0, k1, x0, x1):
    out = (C.c_uint32 * 2)()
    lib().oracle_threefry2x32(C.c_uint32(k0), C.c_uint32(k1), C.c_uint32(x0), C.c_uint32(x1), out)
    return int(out[0]), int(out[1])


def fold_in(key, data):
    k = (C.c_uint32 * 2)(*[int(v) for v in key])
    out = (C.c_uint32 * 2)()
    lib().oracle_fold_in(k, C.c_uint32(int(data) & 0xFFFFFFFF), out)
    return int(out[0]), int(out[1])


def normal(key, n, partitionable=True) -> np.ndarray:
    k = (C.c_uint32 * 2)(*[int(v) for v in key])
    out = np.empty(n, dtype=np.float32)
    lib().oracle_normal(k, C.c_int(n), C.c_int(int(partitionable)), C.c_void_p(out.ctypes.data))
    return out


def tree_eval(tree: np.ndarray, table: abi.NodeTable, data) -> float:
    tree = np.ascontiguousarray(tree, dtype=np.float32)
    data = np.ascontiguousarray(data, dtype=np.float32)
    n = tree.shape[0]
    scratch = np.empty(n + 1, dtype=np.float32)
    return float(lib().oracle_tree_eval(C.c_void_p(tree.ctypes.data), C.c_int(n), C.byref(table),
                                        C.c_void_p(data.ctypes.data), C.c_void_p(scratch.ctypes.data)))


def schedule(ts, dt0, max_steps=16 ** 4):
    ts = np.ascontiguousarray(ts, dtype=np.float32)
    cap = int(max_steps) if max_steps < (1 << 22) else (1 << 22)
    tp = np.empty(cap, dtype=np.float32)
    tn = np.empty(cap, dtype=np.float32)
    n = lib().oracle_schedule(C.c_void_p(ts.ctypes.data), C.c_int(ts.size), C.c_float(dt0), C.c_int(max_steps),
                              C.c_void_p(tp.ctypes.data), C.c_void_p(tn.ctypes.data), C.c_int(cap))
    return tp[:n].copy(), tn[:n].copy()


def rollout_population(pop, table, desc_arrays, n_threads=0, trajectories=False):
    """``desc_arrays`` is a RolloutData (multitreegp_b200.rollout_data) holding numpy arrays.

    Returns dict(fitness, traj_fitness, steps[, xs, ys, us, activities])."""
    pop = np.ascontiguousarray(pop, dtype=np.float32)
    P, T, N, _ = pop.shape
    desc, keep = desc_arrays.to_desc_host()
    B, K, D = desc.B, desc.K, desc.D
    S = desc.state_size if desc.kind == abi.EVAL_DYNAMIC else 0
    nc = 0 if desc.kind == abi.EVAL_SR else T - S
    fitness = np.empty(P, dtype=np.float32)
    traj_fitness = np.empty((P, B), dtype=np.float32)
    outs = {}
    ptrs = [None, None, None, None]
    if trajectories:
        outs["xs"] = np.empty((P, B, K, D), dtype=np.float32)
        ptrs[0] = outs["xs"].ctypes.data
        if desc.kind != abi.EVAL_SR:
            outs["ys"] = np.empty((P, B, K, desc.n_obs), dtype=np.float32)
            outs["us"] = np.empty((P, B, K, nc), dtype=np.float32)
            ptrs[1], ptrs[2] = outs["ys"].ctypes.data, outs["us"].ctypes.data
        if desc.kind == abi.EVAL_DYNAMIC:
            outs["activities"] = np.empty((P, B, K, S), dtype=np.float32)
            ptrs[3] = outs["activities"].ctypes.data
    steps = C.c_ulonglong(0)
    rc = lib().oracle_rollout_population(
        C.c_void_p(pop.ctypes.data), C.c_int(P), C.c_int(T), C.c_int(N), C.byref(table), C.byref(desc),
        C.c_void_p(fitness.ctypes.data), C.c_void_p(traj_fitness.ctypes.data),
        C.c_void_p(ptrs[0]), C.c_void_p(ptrs[1]), C.c_void_p(ptrs[2]), C.c_void_p(ptrs[3]),
        C.byref(steps), C.c_int(n_threads))
    if rc != 0:
        raise ValueError("oracle_rollout_population: unsupported dimensions")
    del keep
    outs.update(fitness=fitness, traj_fitness=traj_fitness, steps=int(steps.value))
    return outs


def pointwise_population(pop, table, X, Y, max_fitness=1e5, n_threads=0, predictions=False):
    pop = np.ascontiguousarray(pop, dtype=np.float32)
    X = np.ascontiguousarray(X, dtype=np.float32)
    Y = np.ascontiguousarray(Y, dtype=np.float32)
    P, T, N, _ = pop.shape
    n_points, n_vars = X.shape
    fitness = np.empty(P, dtype=np.float32)
    pred = np.empty((P, n_points, T), dtype=np.float32) if predictions else None
    lib().oracle_pointwise_population(
        C.c_void_p(pop.ctypes.data), C.c_int(P), C.c_int(T), C.c_int(N), C.byref(table),
        C.c_void_p(X.ctypes.data), C.c_void_p(Y.ctypes.data), C.c_int(n_points), C.c_int(n_vars),
        C.c_float(max_fitness), C.c_void_p(fitness.ctypes.data),
        C.c_void_p(pred.ctypes.data if pred is not None else None), C.c_int(n_threads))
    return (fitness, pred) if predictions else fitness


def live_count(pop) -> np.ndarray:
    pop = np.ascontiguousarray(pop, dtype=np.float32)
    P, T, N, _ = pop.shape
    out = np.empty(P, dtype=np.int32)
    lib().oracle_live_count(C.c_void_p(pop.ctypes.data), C.c_int(P), C.c_int(T), C.c_int(N),
                            C.c_void_p(out.ctypes.data))
    return out


def num_threads() -> int:
    return int(lib().oracle_num_threads())
