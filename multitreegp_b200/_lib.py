"""Loader of libmtgp_b200.so (the CUDA product).  There is NO fallback: if the library is missing
or fails to load, every evaluator raises -- a silent CPU path would void the parity claims."""
from __future__ import annotations

import ctypes as C
import os

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MTGP_LIB") or os.path.join(_HERE, "libmtgp_b200.so")  # MTGP_LIB: tuning builds
_LIB = None


class MtgpError(RuntimeError):
    def __init__(self, code: int, message: str) -> None:
        super().__init__(f"libmtgp_b200 error {code}: {message}")
        self.code = code


def lib() -> C.CDLL:
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m multitreegp_b200.build` "
            "(nvcc, sm_100a).  multitreegp_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, f32, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
    L.mtgp_abi_version.restype = i32
    L.mtgp_last_error.restype = C.c_char_p
    L.mtgp_workspace_bytes.restype = sz
    L.mtgp_workspace_bytes.argtypes = [i32, i32, i32, vp]
    L.mtgp_workspace_bytes_pointwise.restype = sz
    L.mtgp_workspace_bytes_pointwise.argtypes = [i32, i32, i32, i32, i32]
    L.mtgp_bytecode_bytes.restype = sz
    L.mtgp_bytecode_bytes.argtypes = [i32, i32, i32]
    L.mtgp_program_stride.restype = i32
    L.mtgp_program_stride.argtypes = [i32]
    tab = C.POINTER(abi.NodeTable)
    desc = C.POINTER(abi.RolloutDesc)
    traj = C.POINTER(abi.TrajectoryOut)
    L.mtgp_encode.restype = i32
    L.mtgp_encode.argtypes = [vp, i32, vp, i32, i32, i32, i32, tab, abi.Bytecode, vp]
    L.mtgp_decode.restype = i32
    L.mtgp_decode.argtypes = [abi.Bytecode, i32, i32, i32, vp, vp]
    L.mtgp_lower.restype = i32
    L.mtgp_lower.argtypes = [abi.Bytecode, i32, i32, i32, tab, i32, vp, vp, vp]
    L.mtgp_rollout.restype = i32
    L.mtgp_rollout.argtypes = [abi.Bytecode, i32, i32, i32, tab, desc, vp, traj, vp, vp, sz, vp]
    L.mtgp_population_fitness.restype = i32
    L.mtgp_population_fitness.argtypes = [vp, i32, vp, i32, i32, i32, i32, tab, desc, f32, vp, vp, vp, sz, vp]
    L.mtgp_population_fitness_ex.restype = i32
    L.mtgp_population_fitness_ex.argtypes = [vp, i32, vp, i32, i32, i32, i32, tab, desc, f32, vp, traj, vp, vp, sz, vp]
    L.mtgp_population_fitness_host.restype = i32
    L.mtgp_population_fitness_host.argtypes = [vp, i32, i32, i32, tab, desc, f32, vp, vp]
    L.mtgp_eval_pointwise.restype = i32
    L.mtgp_eval_pointwise.argtypes = [abi.Bytecode, i32, i32, i32, tab, vp, vp, i32, i32, f32, vp, vp, vp, sz, vp]
    L.mtgp_population_fitness_pointwise.restype = i32
    L.mtgp_population_fitness_pointwise.argtypes = [vp, i32, vp, i32, i32, i32, i32, tab, vp, vp, i32, i32, f32, f32,
                                                    vp, vp, sz, vp]
    L.mtgp_population_fitness_pointwise_host.restype = i32
    L.mtgp_population_fitness_pointwise_host.argtypes = [vp, i32, i32, i32, tab, vp, vp, i32, i32, f32, f32, vp, vp]
    L.mtgp_select.restype = i32
    L.mtgp_select.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp, i32, vp, vp]
    if L.mtgp_abi_version() != abi.MTGP_ABI_VERSION:
        raise RuntimeError("libmtgp_b200.so ABI version mismatch: rebuild with python -m multitreegp_b200.build")
    _LIB = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        raise MtgpError(rc, lib().mtgp_last_error().decode())


#: every symbol include/mtgp.h declares (tests assert the .so exports them all)
EXPORTED_SYMBOLS = [
    "mtgp_abi_version", "mtgp_last_error", "mtgp_workspace_bytes", "mtgp_workspace_bytes_pointwise",
    "mtgp_bytecode_bytes", "mtgp_program_stride", "mtgp_encode", "mtgp_decode", "mtgp_lower",
    "mtgp_eval_pointwise", "mtgp_rollout", "mtgp_population_fitness", "mtgp_population_fitness_ex",
    "mtgp_population_fitness_host", "mtgp_population_fitness_pointwise", "mtgp_population_fitness_pointwise_host",
    "mtgp_select", "mtgp_release_scratch",
]
