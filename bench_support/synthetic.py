"""Synthetic populations and data sets of the reference's shapes (NumPy, host side).

``sample_population`` is a distribution-equivalent restatement of the reference initialiser
(genetic_operators/initialization.py:9-124 with the breadth-first -> depth-first map of
genetic_programming.py:272-296): breadth-first sampling into a full 2^depth-1 scaffold stored at
depth-first positions, then pruning of the empty rows into a right-aligned (max_nodes, 4) array.
It does not reproduce jax.random streams (JAX is not available here); it reproduces the tree
ARRAY FORMAT and the sampling distribution, vectorised over all trees of the population.
"""
from __future__ import annotations

import numpy as np

from multitreegp_b200.node_library import NodeLibrary


def create_map_b_to_d(depth: int) -> np.ndarray:
    """Breadth-first index -> storage row of the full scaffold (genetic_programming.py:272-296)."""
    max_nodes = 2 ** depth - 1
    current_depth = 0
    m = np.zeros(max_nodes, dtype=np.int64)
    for i in range(max_nodes):
        if i > 0:
            parent = (i + (i % 2) - 2) // 2
            value = m[parent]
            if i % 2 == 0:  # right child
                new_value = value + 2 ** (depth - current_depth + 1)
            else:  # left child
                new_value = value + 1
            m[i] = new_value
        current_depth += int(i == (2 ** current_depth - 1))
    return max_nodes - 1 - m


def sample_population(rng: np.random.Generator, population_size: int, library: NodeLibrary,
                      max_init_depth: int = 4, max_nodes: int = 30, coefficient_sd: float = 1.0,
                      full: bool = False, op_growth: float = 0.7) -> np.ndarray:
    """Returns f32[population_size, num_trees, max_nodes, 4] in the reference's array format.

    ``full=True`` makes every internal position an operator (worst-case 2^depth-1 live rows);
    ``op_growth`` is the base of the reference's operator probability ``0.7 ** depth``
    (initialization.py:35) -- larger values give the bloated trees of an evolved population.
    """
    T = library.num_trees
    M = population_size * T
    tree_size = 2 ** max_init_depth - 1
    assert max_nodes >= 1
    b2d = create_map_b_to_d(max_init_depth)
    slots = library.slots
    ids = np.zeros((M, tree_size), dtype=np.int64)
    a = -np.ones((M, tree_size), dtype=np.int64)
    b = -np.ones((M, tree_size), dtype=np.int64)
    coef = np.zeros((M, tree_size), dtype=np.float32)
    open_slots = np.ones(M, dtype=np.int64)

    op_p = library.operator_probabilities / library.operator_probabilities.sum()
    if full:  # only binary operators: every one of the 2^depth-1 scaffold rows is live
        op_p = np.where(slots[library.operator_indices] == 2, op_p, 0.0)
        op_p = op_p / op_p.sum()
    op_cdf = np.cumsum(op_p)
    var_p = library.variable_array / library.variable_array.sum(axis=1, keepdims=True)
    var_cdf = np.tile(np.cumsum(var_p, axis=1), (population_size, 1))  # [M, n_data]

    for i in range(tree_size):
        _i = b2d[i]
        depth = int(np.floor(np.log2(i + 1) + 1e-9))
        coefficient = (rng.standard_normal(M) * coefficient_sd).astype(np.float32)
        u_var = rng.random(M)
        var_choice = library.variable_indices[np.minimum((u_var[:, None] > var_cdf).sum(axis=1),
                                                         len(library.variable_indices) - 1)]
        leaf = np.where(rng.random(M) < 0.5, 1, var_choice)
        op_choice = library.operator_indices[np.minimum(np.searchsorted(op_cdf, rng.random(M)),
                                                        len(op_cdf) - 1)]
        grow = (open_slots < max_nodes - i - 1) & (depth + 1 < max_init_depth)
        p_op = 1.0 if full else op_growth ** depth
        index = np.where(grow, np.where(rng.random(M) < p_op, op_choice, leaf), leaf)
        index = np.where(open_slots == 0, 0, index)
        if i > 0:
            parent = (i + (i % 2) - 2) // 2
            parent_id = np.maximum(ids[:, b2d[parent]], 0)
            index = np.where((slots[parent_id] + i % 2) > 1, index, 0)
        ar = slots[index]
        if 2 * i + 1 < tree_size:
            a[:, _i] = np.where(ar > 0, b2d[2 * i + 1], -1)
        if 2 * i + 2 < tree_size:
            b[:, _i] = np.where(ar > 1, b2d[2 * i + 2], -1)
        coef[:, _i] = np.where(index == 1, coefficient, 0.0)
        ids[:, _i] = index
        open_slots = np.where(index == 0, open_slots, np.maximum(0, open_slots + ar - 1))

    # prune_tree (initialization.py:56-98): drop empty rows, right-align, renumber references
    keep = ids != 0
    n_keep = keep.sum(axis=1)
    assert n_keep.max() <= max_nodes, "sampled tree exceeds max_nodes"
    # new row of a kept scaffold row r: (max_nodes-1) - (# kept rows with scaffold index > r)
    kept_after = np.cumsum(keep[:, ::-1], axis=1)[:, ::-1] - keep
    new_row = (max_nodes - 1) - kept_after
    out = np.zeros((M, max_nodes, 4), dtype=np.float32)
    out[:, :, 1:3] = -1.0
    rows_m, rows_r = np.nonzero(keep)
    dst = new_row[rows_m, rows_r]
    out[rows_m, dst, 0] = ids[rows_m, rows_r]
    a_old, b_old = a[rows_m, rows_r], b[rows_m, rows_r]
    out[rows_m, dst, 1] = np.where(a_old >= 0, new_row[rows_m, np.maximum(a_old, 0)], -1)
    out[rows_m, dst, 2] = np.where(b_old >= 0, new_row[rows_m, np.maximum(b_old, 0)], -1)
    out[rows_m, dst, 3] = coef[rows_m, rows_r]
    return out.reshape(population_size, T, max_nodes, 4)


def random_keys(rng: np.random.Generator, batch_size: int) -> np.ndarray:
    """Stand-in for jr.split(key, batch_size): [B, 2] uint32 threefry keys."""
    return rng.integers(0, 2 ** 32, size=(batch_size, 2), dtype=np.uint64).astype(np.uint32)


def van_der_pol_data(rng: np.random.Generator, batch_size: int = 16, T: float = 20.0, dt: float = 0.2,
                     mu: float = 1.0):
    """SR data like examples/SymbolicRegression.ipynb:47-76: x0 ~ N(0, I), ts = arange(0, T, dt),
    trajectories of dx0 = x1, dx1 = mu*(1-x0^2)*x1 - x0 (RK4 with 20 sub-steps per save interval,
    float64, instead of the notebook's Dopri5 -- it is only synthetic ground truth)."""
    x0 = rng.standard_normal((batch_size, 2))
    ts = np.arange(0, T, dt, dtype=np.float32)

    def f(x):
        return np.stack([x[:, 1], mu * (1 - x[:, 0] ** 2) * x[:, 1] - x[:, 0]], axis=1)

    xs = np.empty((batch_size, ts.size, 2))
    x = x0.copy()
    xs[:, 0] = x
    sub = 20
    for k in range(1, ts.size):
        h = float(ts[k] - ts[k - 1]) / sub
        for _ in range(sub):
            k1 = f(x); k2 = f(x + 0.5 * h * k1); k3 = f(x + 0.5 * h * k2); k4 = f(x + h * k3)
            x = x + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
        xs[:, k] = x
    return x0.astype(np.float32), ts, xs.astype(np.float32), random_keys(rng, batch_size)


def pointwise_data(rng: np.random.Generator, n_points: int, n_vars: int = 2):
    """C4 data: X ~ U(-3, 3), y from a fixed ground-truth expression."""
    X = rng.uniform(-3.0, 3.0, size=(n_points, n_vars)).astype(np.float32)
    y = X[:, 0] * X[:, 0] + np.sin(2.0 * X[:, 1 % n_vars]) - 0.5 * X[:, 0] * X[:, 1 % n_vars]
    return X, y.astype(np.float32).reshape(n_points, 1)
