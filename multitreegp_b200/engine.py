"""Host-side driver of the CUDA fitness path: torch is used only for device memory and streams.

``FitnessEngine`` owns nothing but a grow-on-demand workspace tensor per device; every evaluation
is encode -> lower -> evaluate enqueued on torch's current stream through the C ABI
(include/mtgp.h).  It mirrors the data flow of ``GeneticProgramming.evaluate_population``
(genetic_programming.py:403-433): flat population f32[P,T,N,4] in, fitness f32[P] out, parsimony
term added on the device.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import abi
from ._lib import check, lib
from .rollout_data import RolloutData


def _ptr(t: torch.Tensor | None):
    if t is None or t.numel() == 0:
        return None
    assert t.is_cuda and t.is_contiguous()
    return t.data_ptr()


class DeviceData:
    """RolloutData resident on a CUDA device."""

    def __init__(self, host: RolloutData, device: torch.device) -> None:
        self.host = host
        self.device = device

        def up(a, dtype=None):
            if a is None:
                return None
            if a.dtype == np.uint32:
                a = a.view(np.int32)
            return torch.from_numpy(np.ascontiguousarray(a)).to(device, non_blocking=True)

        self.x0, self.ts = up(host.x0), up(host.ts)
        self.targets = up(host.targets) if host.targets.size else None
        self.obs_keys, self.params, self.ys_true = up(host.obs_keys), up(host.params), up(host.ys_true)

    def desc(self) -> abi.RolloutDesc:
        lookup = {id(self.host.x0): self.x0, id(self.host.ts): self.ts, id(self.host.targets): self.targets,
                  id(self.host.obs_keys): self.obs_keys, id(self.host.params): self.params,
                  id(self.host.ys_true): self.ys_true}
        return self.host.fill_desc(lambda a: None if a is None else _ptr(lookup[id(a)]))


class FitnessEngine:
    def __init__(self, device: torch.device | str | None = None) -> None:
        if not torch.cuda.is_available():
            raise RuntimeError("multitreegp_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.lib = lib()
        self._ws = None

    # -- workspace -------------------------------------------------------------------------
    def _workspace(self, nbytes: int) -> torch.Tensor:
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(int(nbytes * 1.05) + 1024, dtype=torch.uint8, device=self.device)
        return self._ws

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    @staticmethod
    def _split(pop: torch.Tensor):
        assert pop.is_cuda and pop.dtype == torch.float32 and pop.is_contiguous() and pop.dim() == 4 and pop.shape[-1] == 4
        P, T, N, _ = pop.shape
        return pop.data_ptr(), 4, pop.data_ptr() + 12, 4, P, T, N

    # -- level-1 bytecode ------------------------------------------------------------------
    def alloc_bytecode(self, P: int, T: int, N: int):
        dev = self.device
        arrs = dict(op=torch.empty((P, T, N), dtype=torch.uint8, device=dev),
                    a=torch.empty((P, T, N), dtype=torch.int16, device=dev),
                    b=torch.empty((P, T, N), dtype=torch.int16, device=dev),
                    coef=torch.empty((P, T, N), dtype=torch.float32, device=dev),
                    live=torch.empty((P, T), dtype=torch.int32, device=dev),
                    flags=torch.empty((P, T), dtype=torch.uint8, device=dev))
        bc = abi.Bytecode(*[arrs[k].data_ptr() for k in ("op", "a", "b", "coef", "live", "flags")])
        return bc, arrs

    def encode(self, pop: torch.Tensor, table: abi.NodeTable):
        n, ns, c, cs, P, T, N = self._split(pop)
        bc, arrs = self.alloc_bytecode(P, T, N)
        check(self.lib.mtgp_encode(n, ns, c, cs, P, T, N, C.byref(table), bc, self._stream()))
        return bc, arrs

    def encode_split(self, nodes: torch.Tensor, coefficients: torch.Tensor, table: abi.NodeTable):
        """The split arrays shard_eval hands to the evaluator (genetic_programming.py:261)."""
        P, T, N, _ = nodes.shape
        bc, arrs = self.alloc_bytecode(P, T, N)
        check(self.lib.mtgp_encode(nodes.data_ptr(), 3, coefficients.data_ptr(), 1, P, T, N, C.byref(table), bc,
                                   self._stream()))
        return bc, arrs

    def decode(self, bc: abi.Bytecode, P: int, T: int, N: int) -> torch.Tensor:
        out = torch.empty((P, T, N, 4), dtype=torch.float32, device=self.device)
        check(self.lib.mtgp_decode(bc, P, T, N, out.data_ptr(), self._stream()))
        return out

    def lower(self, bc: abi.Bytecode, P: int, T: int, N: int, table: abi.NodeTable, tmp_base: int):
        ls = self.lib.mtgp_program_stride(N)
        prog = torch.empty((P * T, ls, 2), dtype=torch.int32, device=self.device)
        meta = torch.empty((P * T,), dtype=torch.int32, device=self.device)
        check(self.lib.mtgp_lower(bc, P, T, N, C.byref(table), tmp_base, prog.data_ptr(), meta.data_ptr(),
                                  self._stream()))
        return prog, meta

    # -- rollouts --------------------------------------------------------------------------
    def rollout_fitness(self, pop: torch.Tensor, table: abi.NodeTable, data: DeviceData,
                        size_parsimony: float = 0.0, trajectories: bool = False, count_steps: bool = False):
        n, ns, c, cs, P, T, N = self._split(pop)
        desc = data.desc()
        nbytes = self.lib.mtgp_workspace_bytes(P, T, N, C.byref(desc))
        ws = self._workspace(nbytes)
        fitness = torch.empty((P,), dtype=torch.float32, device=self.device)
        steps = torch.zeros((1,), dtype=torch.int64, device=self.device) if count_steps else None
        outs, traj_p = {}, None
        if trajectories:
            B, K, D = desc.B, desc.K, desc.D
            S = desc.state_size if desc.kind == abi.EVAL_DYNAMIC else 0
            mk = lambda *shape: torch.empty(shape, dtype=torch.float32, device=self.device)
            outs["traj_fitness"] = mk(P, B)
            outs["xs"] = mk(P, B, K, D)
            tr = abi.TrajectoryOut()
            tr.xs, tr.traj_fitness = outs["xs"].data_ptr(), outs["traj_fitness"].data_ptr()
            if desc.kind != abi.EVAL_SR:
                outs["ys"], outs["us"] = mk(P, B, K, desc.n_obs), mk(P, B, K, 1)
                tr.ys, tr.us = outs["ys"].data_ptr(), outs["us"].data_ptr()
            if desc.kind == abi.EVAL_DYNAMIC:
                outs["activities"] = mk(P, B, K, S)
                tr.activities = outs["activities"].data_ptr()
            traj_p = C.byref(tr)
        check(self.lib.mtgp_population_fitness_ex(
            n, ns, c, cs, P, T, N, C.byref(table), C.byref(desc), float(size_parsimony), fitness.data_ptr(),
            traj_p, None if steps is None else steps.data_ptr(), ws.data_ptr(), ws.numel(), self._stream()))
        outs["fitness"] = fitness
        if steps is not None:
            outs["steps"] = steps
        return outs

    def rollout_fitness_host(self, pop: np.ndarray, table: abi.NodeTable, data: RolloutData,
                             size_parsimony: float = 0.0, out: np.ndarray | None = None) -> np.ndarray:
        """Host buffers in, host buffer out: H2D + evaluation + D2H inside one C call."""
        assert pop.dtype == np.float32 and pop.flags["C_CONTIGUOUS"]
        P, T, N, _ = pop.shape
        desc, keep = data.to_desc_host()
        fitness = np.empty(P, dtype=np.float32) if out is None else out
        with torch.cuda.device(self.device):
            check(self.lib.mtgp_population_fitness_host(pop.ctypes.data, P, T, N, C.byref(table), C.byref(desc),
                                                        float(size_parsimony), fitness.ctypes.data, self._stream()))
        del keep
        return fitness

    # -- pointwise -------------------------------------------------------------------------
    def pointwise_fitness(self, pop: torch.Tensor, table: abi.NodeTable, X: torch.Tensor, Y: torch.Tensor,
                          max_fitness: float = 1e5, size_parsimony: float = 0.0, predictions: bool = False):
        n, ns, c, cs, P, T, N = self._split(pop)
        n_points, n_vars = X.shape
        assert Y.shape == (n_points, T) and X.is_contiguous() and Y.is_contiguous()
        nbytes = self.lib.mtgp_workspace_bytes_pointwise(P, T, N, n_points, n_vars)
        ws = self._workspace(nbytes)
        fitness = torch.empty((P,), dtype=torch.float32, device=self.device)
        if predictions:
            bc, arrs = self.encode(pop, table)
            pred = torch.empty((P, n_points, T), dtype=torch.float32, device=self.device)
            check(self.lib.mtgp_eval_pointwise(bc, P, T, N, C.byref(table), X.data_ptr(), Y.data_ptr(), n_points,
                                               n_vars, float(max_fitness), fitness.data_ptr(), pred.data_ptr(),
                                               ws.data_ptr(), ws.numel(), self._stream()))
            if size_parsimony:
                fitness = fitness + size_parsimony * arrs["live"].sum(dim=1).float()
            return fitness, pred
        check(self.lib.mtgp_population_fitness_pointwise(
            n, ns, c, cs, P, T, N, C.byref(table), X.data_ptr(), Y.data_ptr(), n_points, n_vars, float(max_fitness),
            float(size_parsimony), fitness.data_ptr(), ws.data_ptr(), ws.numel(), self._stream()))
        return fitness

    def pointwise_fitness_host(self, pop: np.ndarray, table: abi.NodeTable, X: np.ndarray, Y: np.ndarray,
                               max_fitness: float = 1e5, size_parsimony: float = 0.0) -> np.ndarray:
        P, T, N, _ = pop.shape
        fitness = np.empty(P, dtype=np.float32)
        with torch.cuda.device(self.device):
            check(self.lib.mtgp_population_fitness_pointwise_host(
                pop.ctypes.data, P, T, N, C.byref(table), X.ctypes.data, Y.ctypes.data, X.shape[0], X.shape[1],
                float(max_fitness), float(size_parsimony), fitness.ctypes.data, self._stream()))
        return fitness

    # -- selection (the consumer of the fitness vector) ------------------------------------
    def select(self, fitness: torch.Tensor, num_populations: int, elite_size: int = 0,
               population: torch.Tensor | None = None):
        """argmin / best candidate (genetic_programming.py:426-431) and the per-sub-population elite
        indices (reproduction.py:82) without leaving the device.  ``fitness`` is the flat f32[P]
        vector (or [num_populations, population_size]); ``population`` f32[P,T,N,4] is optional.
        Returns (best_index i32[1], best_fitness f32[1], elite i32[num_populations, elite_size],
        best_solution f32[T,N,4] | None) as device tensors -- nothing is synchronised."""
        f = fitness.reshape(-1)
        assert f.is_cuda and f.dtype == torch.float32 and f.is_contiguous()
        P = f.numel()
        assert P % num_populations == 0
        best_i = torch.empty((1,), dtype=torch.int32, device=self.device)
        best_f = torch.empty((1,), dtype=torch.float32, device=self.device)
        elite = torch.empty((num_populations, elite_size), dtype=torch.int32, device=self.device)
        best_sol, pop_ptr, row = None, None, 0
        if population is not None:
            assert population.is_cuda and population.dtype == torch.float32 and population.is_contiguous()
            assert population.shape[0] == P
            best_sol = torch.empty(tuple(population.shape[1:]), dtype=torch.float32, device=self.device)
            pop_ptr, row = population.data_ptr(), best_sol.numel()
        check(self.lib.mtgp_select(f.data_ptr(), num_populations, P // num_populations, elite_size,
                                   best_i.data_ptr(), best_f.data_ptr(), elite.data_ptr() if elite_size else None,
                                   pop_ptr, row, best_sol.data_ptr() if best_sol is not None else None,
                                   self._stream()))
        return best_i, best_f, elite, best_sol
