"""BASELINE.json's full sizes on the GPU, checked through size-independent properties (the oracle
cannot evaluate 65 536 x 32 x 2 000 rollout-steps in seconds): oracle spot checks on a random
subsample, idempotence, permutation equivariance, shard consistency, parsimony linearity, and
empty / maximum-size inputs."""
import numpy as np
import pytest

from multitreegp_b200 import abi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from multitreegp_b200.engine import FitnessEngine
    return FitnessEngine()


def _dev(engine, a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).to(engine.device)


def test_c5_full_size_properties(engine, oracle):
    """DynamicPolicy sweep: 65 536 individuals x 32 initial conditions x 2 000 Euler steps (bench.py's workload)."""
    import bench
    from multitreegp_b200.engine import DeviceData
    w = bench.make_rollout_workload("c5", seed=3)
    pop, lib, data = w["pop"], w["lib"], w["data"]
    P = pop.shape[0]
    dd = DeviceData(data, engine.device)
    pop_d = _dev(engine, pop)
    out = engine.rollout_fitness(pop_d, lib.table(), dd, count_steps=True)
    fit = out["fitness"].cpu().numpy()
    assert fit.shape == (P,) and np.isfinite(fit).all() and (fit >= 0).all() and (fit <= 1e4).all()
    assert 0 < int(out["steps"].item()) <= P * 32 * w["n_sched"]
    rng = np.random.default_rng(0)
    # 1. oracle on a random subsample of the full-size population
    idx = rng.choice(P, 48, replace=False)
    ref = oracle.rollout_population(pop[idx], lib.table(), data)["fitness"]
    np.testing.assert_allclose(fit[idx], ref, rtol=1e-5)
    # 2. idempotence: the same launch twice gives the same bits
    again = engine.rollout_fitness(pop_d, lib.table(), dd)["fitness"].cpu().numpy()
    assert (again.view(np.uint32) == fit.view(np.uint32)).all()
    # 3. permutation equivariance: individuals are independent
    perm = rng.permutation(P)
    fit_p = engine.rollout_fitness(_dev(engine, pop[perm]), lib.table(), dd)["fitness"].cpu().numpy()
    assert (fit_p.view(np.uint32) == fit[perm].view(np.uint32)).all()
    # 4. shard consistency: what the 8-GPU run computes per rank (contiguous eighths) equals the whole
    lo, hi = 3 * (P // 8), 4 * (P // 8)
    shard = engine.rollout_fitness(_dev(engine, pop[lo:hi]), lib.table(), dd)["fitness"].cpu().numpy()
    assert (shard.view(np.uint32) == fit[lo:hi].view(np.uint32)).all()
    # 5. parsimony is linear in the live-row count (genetic_programming.py:424)
    with_p = engine.rollout_fitness(pop_d, lib.table(), dd, size_parsimony=0.25)["fitness"].cpu().numpy()
    live = (pop[..., 0] != 0).sum(axis=(1, 2)).astype(np.float32)
    np.testing.assert_array_equal(with_p, fit + np.float32(0.25) * live)


def test_c4_full_size_properties(engine, oracle):
    """Pointwise scale-up: 100 000 individuals x 100 000 points (max_nodes 255)."""
    import bench
    from multitreegp_b200.node_library import NodeLibrary
    from bench_support.synthetic import pointwise_data, sample_population
    rng = np.random.default_rng(7)
    c = bench.C4
    lib = NodeLibrary(bench.OPS_SR, [[f"x{i}" for i in range(c["n_vars"])]], [1])
    pop = sample_population(rng, c["P"], lib, c["depth"], c["N"])
    X, Y = pointwise_data(rng, c["n_points"], c["n_vars"])
    Xd, Yd = _dev(engine, X), _dev(engine, Y)
    fit = engine.pointwise_fitness(_dev(engine, pop), lib.table(), Xd, Yd).cpu().numpy()
    assert np.isfinite(fit).all() and (fit >= 0).all() and (fit <= 1e5).all()
    idx = rng.choice(c["P"], 24, replace=False)
    ref = oracle.pointwise_population(pop[idx], lib.table(), X, Y)
    np.testing.assert_allclose(fit[idx], ref, rtol=1e-5)
    perm = rng.permutation(c["P"])
    fit_p = engine.pointwise_fitness(_dev(engine, pop[perm]), lib.table(), Xd, Yd).cpu().numpy()
    assert (fit_p.view(np.uint32) == fit[perm].view(np.uint32)).all()
    # a constant tree y = c has MSE mean((c - Y)^2): closed form over all 100 000 points
    const = np.tile(np.float32([0, -1, -1, 0]), (c["N"], 1))
    const[-1] = [1, -1, -1, 0.5]
    f_const = engine.pointwise_fitness(_dev(engine, const[None, None]), lib.table(), Xd, Yd).cpu().numpy()[0]
    np.testing.assert_allclose(f_const, np.mean((0.5 - Y.astype(np.float64)) ** 2), rtol=1e-5)


def test_empty_and_maximum_sizes(engine, oracle):
    from multitreegp_b200.engine import DeviceData
    from multitreegp_b200.environments.control_environments import HarmonicOscillator
    from bench_support.synthetic import sample_population
    from tests import util
    import torch
    rng = np.random.default_rng(5)
    env = HarmonicOscillator(0.0, 0.0)
    lib = util.control_library(env, abi.EVAL_STATIC, ops=util.OPS_SR)
    data = util.control_data(env, abi.EVAL_STATIC, rng, 4, 2.0, 0.2, 0.05, 0, "Different")
    dd = DeviceData(data, engine.device)
    # empty population: a legal no-op
    empty = torch.empty((0, 1, 30, 4), dtype=torch.float32, device=engine.device)
    assert engine.rollout_fitness(empty, lib.table(), dd)["fitness"].shape == (0,)
    # the largest trees the value file can hold: max_nodes 400, every row of a depth-8 scaffold live
    pop = sample_population(rng, 6, lib, 8, 400, full=True)
    assert ((pop[..., 0] != 0).sum(-1) == 255).all()
    ref = oracle.rollout_population(pop, lib.table(), data)
    out = engine.rollout_fitness(_dev(engine, pop), lib.table(), dd, trajectories=True)
    assert util.same_or_both_nan(out["traj_fitness"].cpu().numpy(), ref["traj_fitness"]).all()
    # beyond the limits: refused with a message, never a wrong answer
    from multitreegp_b200._lib import MtgpError
    with pytest.raises(MtgpError):
        engine.rollout_fitness(_dev(engine, np.zeros((2, 1, 600, 4), np.float32)), lib.table(), dd)


def test_host_entry_overlapped_upload_equals_device_entry(engine):
    """mtgp_population_fitness_host uploads populations of 32 MB and more as a prologue slice + the rest, evaluated on two
    streams while the upload is still running: the fitness vector must be the device entry point's, bit for bit
    (and a population below the threshold, which takes the single-upload path, likewise)."""
    from multitreegp_b200.engine import DeviceData
    from tests import util
    rng = np.random.default_rng(17)
    for P in (12000, 2000):  # 3 trees x 63 rows x 16 B = 3 024 B per individual: 36 MB / 6 MB
        lib, pop, data = util.control_case("acrobot", abi.EVAL_DYNAMIC, rng, P=P, B=8, T_end=1.0, dt0=0.05, obs_noise=0.1,
                                           depth=6, N=63)
        assert (pop.nbytes >= (32 << 20)) == (P == 12000)
        dev = engine.rollout_fitness(_dev(engine, pop), lib.table(), DeviceData(data, engine.device),
                                     size_parsimony=0.01)["fitness"].cpu().numpy()
        host = engine.rollout_fitness_host(pop, lib.table(), data, size_parsimony=0.01)
        assert (host.view(np.uint32) == dev.view(np.uint32)).all()
        again = engine.rollout_fitness_host(pop, lib.table(), data, size_parsimony=0.01)  # streams / events reused
        assert (again.view(np.uint32) == dev.view(np.uint32)).all()


def test_pointwise_host_entry_overlapped_upload_equals_device_entry(engine):
    """mtgp_population_fitness_pointwise_host with a population above the 32 MB threshold (four slices on two streams,
    a ragged last slice) and below it: the bits of the device entry point."""
    import bench
    from multitreegp_b200.node_library import NodeLibrary
    from bench_support.synthetic import pointwise_data, sample_population
    rng = np.random.default_rng(23)
    lib = NodeLibrary(bench.OPS_SR, [["x0", "x1"]], [1])
    X, Y = pointwise_data(rng, 700, 2)
    for P in (9001, 500):  # 255 rows x 16 B = 4 080 B per individual: 36.7 MB / 2 MB
        pop = sample_population(rng, P, lib, 6, 255)
        assert (pop.nbytes >= (32 << 20)) == (P == 9001)
        dev = engine.pointwise_fitness(_dev(engine, pop), lib.table(), _dev(engine, X), _dev(engine, Y),
                                       size_parsimony=0.002).cpu().numpy()
        for _ in range(2):  # streams / events / workspace halves reused
            host = engine.pointwise_fitness_host(pop, lib.table(), X, Y, size_parsimony=0.002)
            assert (host.view(np.uint32) == dev.view(np.uint32)).all()
