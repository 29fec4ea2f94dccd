"""Two ranks, two GPUs, NCCL: the population-sharded evaluation (the reference's shard_map over the
population axis, genetic_programming.py:255-262, :412-415) must give every rank the SAME fitness vector as
one GPU evaluating the whole population -- bit for bit, in the original order."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from multitreegp_b200 import abi
    from multitreegp_b200.distributed import evaluate_population_sharded
    from multitreegp_b200.engine import DeviceData, FitnessEngine
    from tests import util
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=dev)
    rng = np.random.default_rng(321)                      # same population and data on every rank
    lib, pop, data = util.control_case("acrobot", abi.EVAL_DYNAMIC, rng, P=256, B=32, T_end=6.0, dt0=0.05, obs_noise=0.1)
    eng = FitnessEngine(dev)
    data_d = DeviceData(data, dev)
    pop_d = torch.from_numpy(pop).to(dev)

    def evaluate_shard(p, d, parsimony):
        return eng.rollout_fitness(p, lib.table(), d, size_parsimony=parsimony)["fitness"]

    fit = evaluate_population_sharded(pop_d, data_d, evaluate_shard, size_parsimony=0.5)
    torch.cuda.synchronize(dev)
    np.save(os.path.join(out_dir, f"fit{rank}.npy"), fit.cpu().numpy())
    if rank == 0:
        np.save(os.path.join(out_dir, "single.npy"), evaluate_shard(pop_d, data_d, 0.5).cpu().numpy())
    dist.destroy_process_group()


@pytest.mark.gpu
def test_nccl_gathered_fitness_equals_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    single = np.load(tmp_path / "single.npy")
    for r in range(2):
        got = np.load(tmp_path / f"fit{r}.npy")
        assert got.shape == single.shape
        assert (got.view(np.uint32) == single.view(np.uint32)).all(), f"rank {r}: gathered vector differs from the single-GPU one"
