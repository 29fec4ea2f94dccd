"""ctypes mirror of include/mtgp.h (the C ABI of libmtgp_b200.so).

Pure declarations: structs, enums and helpers to fill them from numpy / torch buffers.  Used by
the product loader (``_lib.py``); the test-side CPU oracle deliberately uses the same struct layouts
and imports them from here (the dependency points from the test infrastructure to the ABI, never back).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

MTGP_ABI_VERSION = 3
MTGP_MAX_NODE_IDS = 64
MTGP_MAX_NODES = 512

# mtgp_opcode
OP_EMPTY, OP_COEF, OP_VAR = 0, 1, 2
OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_MAX, OP_MIN = 3, 4, 5, 6, 7, 8
OP_SIN, OP_COS, OP_EXP, OP_LOG, OP_SQRT, OP_TANH, OP_NEG, OP_ABS, OP_SQUARE = 9, 10, 11, 12, 13, 14, 15, 16, 17

#: display string (what the reference keys its operators by, genetic_programming.py:143-162) -> (opcode, arity)
OPERATOR_STRINGS = {
    "+": (OP_ADD, 2), "-": (OP_SUB, 2), "*": (OP_MUL, 2), "/": (OP_DIV, 2),
    "max": (OP_MAX, 2), "min": (OP_MIN, 2), "maximum": (OP_MAX, 2), "minimum": (OP_MIN, 2),
    "sin": (OP_SIN, 1), "cos": (OP_COS, 1), "exp": (OP_EXP, 1), "log": (OP_LOG, 1),
    "sqrt": (OP_SQRT, 1), "tanh": (OP_TANH, 1), "neg": (OP_NEG, 1), "abs": (OP_ABS, 1),
    "square": (OP_SQUARE, 1),
}

# mtgp_env_id / mtgp_evaluator_kind
ENV_NONE, ENV_ACROBOT, ENV_HARMONIC, ENV_REACTOR = 0, 1, 2, 3
EVAL_SR, EVAL_STATIC, EVAL_DYNAMIC = 0, 1, 2
SOLVER_EULER, SOLVER_DOPRI5 = 0, 1
CONTROLLER_CONSTANT, CONTROLLER_PID = 0, 1

BC_CANONICAL, BC_LOSSY = 1, 2

# mtgp_status
OK, ERR_INVALID_ARGUMENT, ERR_UNSUPPORTED, ERR_WORKSPACE, ERR_CUDA = 0, -1, -2, -3, -4


class NodeTable(C.Structure):
    _fields_ = [
        ("n_ids", C.c_int32),
        ("opcode", C.c_uint8 * MTGP_MAX_NODE_IDS),
        ("arg", C.c_int32 * MTGP_MAX_NODE_IDS),
    ]


class Bytecode(C.Structure):
    _fields_ = [
        ("op", C.c_void_p),
        ("a", C.c_void_p),
        ("b", C.c_void_p),
        ("coef", C.c_void_p),
        ("live", C.c_void_p),
        ("flags", C.c_void_p),
    ]


class RolloutDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("env_id", C.c_int32),
        ("B", C.c_int32),
        ("K", C.c_int32),
        ("D", C.c_int32),
        ("n_obs", C.c_int32),
        ("n_targets", C.c_int32),
        ("n_params", C.c_int32),
        ("state_size", C.c_int32),
        ("max_steps", C.c_int32),
        ("dt0", C.c_float),
        ("max_fitness", C.c_float),
        ("obs_noise", C.c_float),
        ("threefry_partitionable", C.c_int32),
        ("solver", C.c_int32),
        ("controller", C.c_int32),
        ("rtol", C.c_float),
        ("atol", C.c_float),
        ("dtmin", C.c_float),
        ("pid_safety", C.c_float),
        ("pid_factormin", C.c_float),
        ("pid_factormax", C.c_float),
        ("pcoeff", C.c_float),
        ("icoeff", C.c_float),
        ("dcoeff", C.c_float),
        ("flags", C.c_uint32),
        ("x0", C.c_void_p),
        ("ts", C.c_void_p),
        ("targets", C.c_void_p),
        ("obs_keys", C.c_void_p),
        ("params", C.c_void_p),
        ("ys_true", C.c_void_p),
        ("params_host", C.c_void_p),
    ]

DESC_NOISE_INLINE = 1


class TrajectoryOut(C.Structure):
    _fields_ = [
        ("xs", C.c_void_p),
        ("ys", C.c_void_p),
        ("us", C.c_void_p),
        ("activities", C.c_void_p),
        ("traj_fitness", C.c_void_p),
    ]


def np_ptr(a: np.ndarray | None):
    """Address of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data if a.size else None
