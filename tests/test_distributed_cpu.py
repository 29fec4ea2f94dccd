"""World-size-2 run of the population-sharding path on CPU (gloo).  The per-shard evaluator is the
CPU oracle here (no GPU in this container); on the GPU box the same function is driven by
bench.py --gpus N with the CUDA evaluator and NCCL."""
import os
import socket
import sys

import numpy as np
import pytest

from multitreegp_b200.distributed import shard_bounds

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds():
    assert [shard_bounds(8, 4, r) for r in range(4)] == [(0, 2), (2, 4), (4, 6), (6, 8)]
    assert shard_bounds(5, 1, 0) == (0, 5)
    with pytest.raises(ValueError):
        shard_bounds(10, 4, 0)          # shard_map needs P % n == 0
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from multitreegp_b200 import abi
    from multitreegp_b200.distributed import evaluate_population_sharded
    from oracle import pyoracle as po
    from tests import util
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rng = np.random.default_rng(123)                      # same population and data on every rank
    lib, pop, data = util.control_case("harmonic", abi.EVAL_DYNAMIC, rng, P=24, B=4, T_end=3.0)

    def evaluate_shard(p, d, parsimony):
        return po.rollout_population(p, lib.table(), d, n_threads=2)["fitness"] + np.float32(parsimony) * po.live_count(p)

    fit = evaluate_population_sharded(pop, data, evaluate_shard, size_parsimony=0.25)
    np.save(os.path.join(out_dir, f"fit{rank}.npy"), fit)
    if rank == 0:
        np.save(os.path.join(out_dir, "ref.npy"), evaluate_shard(pop, data, 0.25))
    try:
        evaluate_population_sharded(pop[:23], data, evaluate_shard)
        ok = False
    except ValueError:
        ok = True
    np.save(os.path.join(out_dir, f"rejects{rank}.npy"), np.array([ok]))
    dist.destroy_process_group()


def test_two_rank_gloo_all_gather(tmp_path):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    ref = np.load(tmp_path / "ref.npy")
    for r in range(2):
        np.testing.assert_array_equal(np.load(tmp_path / f"fit{r}.npy"), ref)   # every rank holds the full vector
        assert np.load(tmp_path / f"rejects{r}.npy")[0]
