"""The product's lowering pass (csrc/mtgp_lower.cuh) and interpreter (csrc/mtgp_program.cuh) compiled
for the HOST by tests/hostlib/host_harness.cu, against the oracle's literal body_fun restatement.
Checks program LOGIC without a GPU; the device build of the same code is checked by -m gpu tests."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from multitreegp_b200.node_library import NodeLibrary
from bench_support.synthetic import sample_population
from tests import util

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hostlib():
    d = os.path.join(HERE, "hostlib")
    so, src = os.path.join(d, "libmtgp_hosttest.so"), os.path.join(d, "host_harness.cu")
    deps = [src] + [os.path.join(HERE, "..", "multitreegp_b200", "csrc", f) for f in ("mtgp_lower.cuh", "mtgp_program.cuh", "mtgp_threaded.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(p) > os.path.getmtime(so) for p in deps):
        subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-fmad=false", "-Xcompiler", "-fPIC", "-Xcompiler",
                               "-ffp-contract=off", "-Xcompiler", "-w", "-DMTGP_TH_FUSE=1", "-shared", "-o", so, src])
    L = C.CDLL(so)
    L.hostlib_eval_tree.restype = C.c_int
    return L


def host_eval(L, tree, tab, data):
    tree = np.ascontiguousarray(tree, np.float32)
    data = np.ascontiguousarray(data, np.float32)
    n_pts, n_data = data.shape
    out = np.empty(n_pts, np.float32)
    out_wide = np.empty(n_pts, np.float32)
    fl = C.c_int(0)
    prog = np.zeros(tree.shape[0] + 4, np.uint64)
    meta = L.hostlib_eval_tree(C.c_void_p(tree.ctypes.data), C.c_int(tree.shape[0]), C.byref(tab),
                               C.c_void_p(data.ctypes.data), C.c_int(n_pts), C.c_int(n_data),
                               C.c_void_p(out.ctypes.data), C.byref(fl), C.c_void_p(prog.ctypes.data), C.c_int(prog.size),
                               C.c_void_p(out_wide.ctypes.data))
    # the compact interpreter (pointwise kernel) and the pre-decoded one (rollout kernel) must agree bit for bit
    assert util.same_or_both_nan(out, out_wide).all(), (out, out_wide)
    return out, meta & 0x3FF, (meta >> 22) & 0x3FF, fl.value, prog


DATA = np.float32([[0.3, -1.2, 2.5], [1.5, 0.25, -0.75], [0, 0, 0], [-2.0, 3.0, 1e-3], [np.inf, 1, -1],
                   [np.nan, 0, 1], [1e20, -1e20, 1e-20]])


def _check(L, oracle, tree, tab):
    got, n_instr, n_tmp, flags, _ = host_eval(L, tree, tab, DATA)
    ref = np.float32([oracle.tree_eval(tree, tab, d) for d in DATA])
    assert util.same_or_both_nan(got, ref).all(), (got, ref)
    return n_instr, n_tmp, flags


@pytest.mark.parametrize("depth,N", [(4, 30), (6, 63), (8, 255)])
def test_random_trees_bit_exact(hostlib, oracle, depth, N):
    rng = np.random.default_rng(depth)
    lib = NodeLibrary(util.OPS_ALL, [["x0", "x1", "x2"]], [1])
    pop = sample_population(rng, 400, lib, depth, N)
    n_ops = lib.slots[pop[:, 0, :, 0].astype(int)] > 0
    for p in range(pop.shape[0]):
        n_instr, n_tmp, flags = _check(hostlib, oracle, pop[p, 0], lib.table())
        assert flags == 1                                  # canonical, not lossy
        assert n_instr <= max(1, n_ops[p].sum())           # one instruction per operator node at most
        assert n_tmp <= depth                              # temporaries bounded by the tree depth


def test_full_trees(hostlib, oracle):
    rng = np.random.default_rng(0)
    lib = NodeLibrary(util.OPS_SR, [["x0", "x1", "x2"]], [1])
    pop = sample_population(rng, 20, lib, 7, 127, full=True)
    assert ((pop[..., 0] != 0).sum(-1) == 127).all()
    for p in range(pop.shape[0]):
        _check(hostlib, oracle, pop[p, 0], lib.table())


def test_non_canonical_and_hostile_trees(hostlib, oracle):
    """Arrays the reference's operators never produce but body_fun still defines: shared sub-trees
    (DAG), forward and self references, a = -1 on operators, out-of-range ids, non-integral fields."""
    lib = NodeLibrary(util.OPS_ALL, [["x0", "x1", "x2"]], [1])
    ids, tab = lib.string_to_node, lib.table()
    E = [0, -1, -1, 0]
    trees = []
    # DAG: row 5 = x0*x1 used three times
    trees.append([E, E, E, [ids["x1"], -1, -1, 0], [ids["x0"], -1, -1, 0], [ids["*"], 4, 3, 0],
                  [ids["+"], 5, 5, 0], [ids["-"], 6, 5, 0]])
    # square of the accumulator and of a shared row
    trees.append([E, E, E, E, [ids["x2"], -1, -1, 0], [ids["sin"], 4, -1, 0], [ids["square"], 5, -1, 0],
                  [ids["*"], 6, 5, 0]])
    # forward reference (row 2 reads row 6's initial coefficient 4.5) and self reference
    trees.append([E, E, [ids["+"], 6, 3, 0], [ids["x0"], -1, -1, 0], E, E, [1, -1, -1, 4.5], [ids["*"], 2, 7, 2.0]])
    # a = -1 on an operator: reads the last row's current column 3
    trees.append([E, E, E, E, E, E, [ids["x1"], -1, -1, 0], [ids["+"], -1, 6, 3.25]])
    # empty rows in the middle of the live block, gaps, operand order reversed
    trees.append([[ids["x0"], -1, -1, 0], E, [1, -1, -1, 2.0], E, [ids["/"], 0, 2, 0], E, E, [ids["neg"], 4, -1, 0]])
    # out-of-range / non-integral / negative node ids and indices
    trees.append([E, E, E, [99, -1, -1, 0], [1.5, -1, -1, 7.0], [-3, -1, -1, 1.0], [ids["+"], 3.9, 4.2, 0],
                  [ids["max"], 6, 5, 0]])
    trees.append([E, E, E, E, E, [ids["x2"], -1, -1, 0], [ids["exp"], 500, -1, 0], [ids["min"], 6, -9, 0]])
    # every binary op with every operand-kind combination (acc/mem/imm on both sides)
    for op in ("+", "-", "*", "/", "max", "min"):
        o = ids[op]
        trees.append([[ids["x1"], -1, -1, 0], [1, -1, -1, -1.5], [ids["x0"], -1, -1, 0],
                      [o, 2, 1, 0],          # mem op imm
                      [o, 1, 0, 0],          # imm op mem
                      [o, 3, 4, 0],          # tmp op acc
                      [o, 5, 1, 0],          # acc op imm
                      [o, 0, 6, 0]])         # mem op acc
    for t in trees:
        _check(hostlib, oracle, np.float32(t), tab)


def test_comb_needs_many_temporaries(hostlib, oracle):
    lib = NodeLibrary(util.OPS_SR, [["x0", "x1", "x2"]], [1])
    ids = lib.string_to_node
    rows = []

    def build(d):                    # right operand (an operator) is computed first and must stay live
        if d == 0:                   # while the whole left sub-tree is evaluated: d temporaries
            rows.append([ids["x0"], -1, -1, 0]); return len(rows) - 1
        rows.append([ids["x1"], -1, -1, 0]); rows.append([1, -1, -1, 0.5 + d])
        rows.append([ids["*"], len(rows) - 1, len(rows) - 2, 0])
        r = len(rows) - 1
        l = build(d - 1)
        rows.append([ids["+"], l, r, 0])
        return len(rows) - 1
    build(10)
    n_instr, n_tmp, flags = _check(hostlib, oracle, np.float32(rows), lib.table())
    assert flags == 1 and n_tmp == 9 and n_instr == 20   # the innermost right operand stays in the accumulator

    rows = []

    def build2(d):                   # both children are operator sub-trees at every level
        if d == 0:
            rows.append([ids["x1"], -1, -1, 0]); rows.append([ids["x0"], -1, -1, 0])
            rows.append([ids["*"], len(rows) - 1, len(rows) - 2, 0]); return len(rows) - 1
        r = build2(d - 1)
        l = build2(d - 1)
        rows.append([ids["+"], l, r, 0])
        return len(rows) - 1
    build2(5)
    n_instr, n_tmp, flags = _check(hostlib, oracle, np.float32(rows), lib.table())
    assert flags == 1 and n_tmp == 5


def test_fuzz_arbitrary_row_arrays(hostlib, oracle):
    """Rows drawn at random -- arbitrary operand indices (backward, forward, self, -1, out of range),
    shared rows, empty rows anywhere.  body_fun defines a value for every such array; the lowering pass
    (reachability, reference counts, spills, constant folding) must reproduce it bit for bit."""
    rng = np.random.default_rng(99)
    lib = NodeLibrary(util.OPS_ALL, [["x0", "x1", "x2"]], [1])
    tab, F = lib.table(), lib.n_ids
    n_checked = 0
    for trial in range(1500):
        N = int(rng.integers(1, 24))
        tree = np.zeros((N, 4), np.float32)
        tree[:, 0] = rng.integers(0, F, N)
        if trial % 3 == 0:   # mostly backward references: deep DAGs with heavy sharing
            tree[:, 1] = [rng.integers(-1, max(i, 1)) for i in range(N)]
            tree[:, 2] = [rng.integers(-1, max(i, 1)) for i in range(N)]
        else:                # anything goes
            tree[:, 1] = rng.integers(-2, N + 2, N)
            tree[:, 2] = rng.integers(-2, N + 2, N)
        tree[:, 3] = np.where(rng.random(N) < 0.7, rng.normal(size=N), 0).astype(np.float32)
        if trial % 7 == 0:
            tree[rng.integers(0, N), 0] = rng.choice([-1.0, 1.5, F + 3, 0.5])
        _check(hostlib, oracle, tree, tab)
        n_checked += 1
    assert n_checked == 1500
