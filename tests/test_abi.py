"""The C-ABI library without a GPU: it loads, exports every symbol include/mtgp.h declares, reports
its version, sizes workspaces, and rejects bad arguments with a message before touching CUDA."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from multitreegp_b200 import _lib, abi
from multitreegp_b200.node_library import NodeLibrary
from tests import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    from multitreegp_b200.build import build_library
    build_library()
    return _lib.lib()


def test_exports_match_header(L):
    header = open(os.path.join(ROOT, "include", "mtgp.h")).read()
    declared = sorted(set(re.findall(r"MTGP_API\s+[\w\s\*]+?\b(mtgp_\w+)\s*\(", header)))
    assert declared == sorted(_lib.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/mtgp.h but not exported"
    assert L.mtgp_abi_version() == abi.MTGP_ABI_VERSION


def test_struct_layouts_match_header():
    assert C.sizeof(abi.NodeTable) == 4 + 64 + 4 * 64
    assert C.sizeof(abi.Bytecode) == 6 * 8
    assert C.sizeof(abi.RolloutDesc) == 26 * 4 + 7 * 8   # + params_host (ABI 3)
    assert C.sizeof(abi.TrajectoryOut) == 5 * 8


def test_workspace_sizes(L):
    small, big = L.mtgp_workspace_bytes(16, 3, 30, None), L.mtgp_workspace_bytes(65536, 3, 63, None)
    assert 0 < small < big < 1 << 30
    assert L.mtgp_bytecode_bytes(10, 2, 30) == 10 * 2 * 30 * 9 + 10 * 2 * 5
    assert L.mtgp_program_stride(63) % 2 == 0 and L.mtgp_program_stride(63) >= 65
    rng = np.random.default_rng(0)
    lib, pop, data = util.control_case("acrobot", abi.EVAL_DYNAMIC, rng, P=4, obs_noise=0.1)
    desc, keep = data.to_desc_host()
    with_noise = L.mtgp_workspace_bytes(4, 3, 30, C.byref(desc))
    assert with_noise > L.mtgp_workspace_bytes(4, 3, 30, None)       # room for the noise table


def _err(L):
    return L.mtgp_last_error().decode()


def test_argument_validation_needs_no_gpu(L):
    lib = NodeLibrary(util.OPS_CONTROL, [["y1", "y2", "y3", "y4"]], [1])
    tab = lib.table()
    bc = abi.Bytecode(1, 1, 1, 1, 1, 1)
    # shapes
    assert L.mtgp_encode(1, 4, 1, 4, -1, 1, 30, C.byref(tab), bc, None) == abi.ERR_INVALID_ARGUMENT
    assert L.mtgp_encode(1, 4, 1, 4, 4, 1, 100000, C.byref(tab), bc, None) == abi.ERR_UNSUPPORTED
    assert "MTGP_MAX_NODES" in _err(L)
    assert L.mtgp_encode(None, 4, 1, 4, 4, 1, 30, C.byref(tab), bc, None) == abi.ERR_INVALID_ARGUMENT
    # node table: id 0 / 1 are reserved, opcodes must be known
    bad = lib.table(); bad.opcode[1] = abi.OP_ADD
    assert L.mtgp_encode(1, 4, 1, 4, 4, 1, 30, C.byref(bad), bc, None) == abi.ERR_INVALID_ARGUMENT
    bad = lib.table(); bad.opcode[3] = 77
    assert L.mtgp_encode(1, 4, 1, 4, 4, 1, 30, C.byref(bad), bc, None) == abi.ERR_INVALID_ARGUMENT
    assert "unknown opcode" in _err(L)
    # rollout descriptor
    rng = np.random.default_rng(1)
    _, _, data = util.control_case("acrobot", abi.EVAL_STATIC, rng, P=4)
    desc, keep = data.to_desc_host()
    args = lambda d, T=1, ws=1 << 20: (1, 4, 1, 4, 4, T, 30, C.byref(tab), C.byref(d), 0.0, 1, None, 1, ws, None)
    assert L.mtgp_population_fitness(*args(desc, ws=16)) == abi.ERR_WORKSPACE
    assert "workspace too small" in _err(L)
    assert L.mtgp_population_fitness(*args(desc, T=2)) == abi.ERR_INVALID_ARGUMENT       # n_control must be 1
    d2 = abi.RolloutDesc.from_buffer_copy(desc); d2.env_id = 9
    assert L.mtgp_population_fitness(*args(d2)) == abi.ERR_INVALID_ARGUMENT
    d2 = abi.RolloutDesc.from_buffer_copy(desc); d2.n_targets = 1
    assert L.mtgp_population_fitness(*args(d2)) == abi.ERR_INVALID_ARGUMENT
    d2 = abi.RolloutDesc.from_buffer_copy(desc); d2.dt0 = 0.0
    assert L.mtgp_population_fitness(*args(d2)) == abi.ERR_INVALID_ARGUMENT
    d2 = abi.RolloutDesc.from_buffer_copy(desc); d2.K = 1
    assert L.mtgp_population_fitness(*args(d2)) == abi.ERR_INVALID_ARGUMENT
    # a variable that reads past the data vector (the reference would fail at trace time)
    wide = NodeLibrary(util.OPS_CONTROL, [["y1", "y2", "y3", "y4", "y5"]], [1]).table()
    a = list(args(desc)); a[7] = C.byref(wide)
    assert L.mtgp_population_fitness(*a) == abi.ERR_INVALID_ARGUMENT
    assert "data vector" in _err(L)


def test_node_library_matches_reference_id_assignment():
    ops = [("+", lambda x, y: x + y, 2, 0.5), ("-", lambda x, y: x - y, 2, 0.1), ("*", lambda x, y: x * y, 2, 0.5),
           ("sin", lambda x: x, 1, 0.1), ("cos", lambda x: x, 1, 0.1), ("+", lambda x, y: x + y, 2, 0.5)]
    lib = NodeLibrary(ops, [["y1", "y2", "y3", "y4", "a1", "a2", "u"], ["a1", "a2"]], [2, 1])
    # genetic_programming.py:143-180: operators de-duplicated by string from id 2, then new variables in order
    assert lib.string_to_node == {"+": 2, "-": 3, "*": 4, "sin": 5, "cos": 6, "y1": 7, "y2": 8, "y3": 9, "y4": 10,
                                  "a1": 11, "a2": 12, "u": 13}
    assert list(lib.slots) == [0, 0, 2, 2, 2, 1, 1, 0, 0, 0, 0, 0, 0, 0]
    assert list(lib.args[7:]) == [0, 1, 2, 3, 4, 5, 6]
    np.testing.assert_array_equal(lib.variable_array, [[1] * 7, [1] * 7, [0, 0, 0, 0, 1, 1, 0]])
    with pytest.raises(ValueError):
        NodeLibrary([("pow", lambda x, y: x ** y, 2)], [["x"]], [1])
    with pytest.raises(ValueError):
        NodeLibrary([("sin", lambda x, y: x, 2)], [["x"]], [1])


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: no product source may import, include or load it."""
    pkg = os.path.join(ROOT, "multitreegp_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".h", ".cc")):
                continue
            src = open(os.path.join(d, f)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"
            assert not re.search(r"#include\s+[\"<][^\">]*oracle", src), f"{f} includes oracle sources"
            assert "libmtgp_oracle" not in src and "pyoracle" not in src, f"{f} loads the oracle"


def test_evaluation_driver_elite_numpy_path():
    """EvaluationDriver.elite on host arrays is the reference's argsort(fitness)[:elite_size] per sub-population
    (reproduction.py:82), stable; the device path (mtgp_select) is compared with it in the GPU suite."""
    from multitreegp_b200.strategy_eval import EvaluationDriver
    drv = EvaluationDriver(1, 4, fitness_function=None, num_populations=2)
    f = np.array([[3.0, 1.0, 1.0, -0.0], [np.nan, 2.0, 0.0, 2.0]], np.float32)
    np.testing.assert_array_equal(drv.elite(f, 3), [[3, 1, 2], [2, 1, 3]])
    assert drv.elite(f, 3).dtype == np.int32
