"""The ``data`` tuple of the reference evaluators as one typed object.

Reference layouts (vmapped ``in_axes`` show which members carry the trajectory axis B):
  SR       (x0 [B,D], ts [K], ys [B,K,D], noise_keys [B,2])                     SR_evaluator.py:55
  control  (x0 [B,D], ts [K], targets [B,n_t], process_keys [B,2], obs_keys [B,2], params pytree of [B])
                                                       feedforward_evaluate.py:62, dynamic_evaluate.py:63
Process-noise keys are dead inputs (only ODETerm(_drift) is solved) and are dropped.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import abi


@dataclass
class RolloutData:
    kind: int
    env_id: int
    x0: np.ndarray                  # [B, D]
    ts: np.ndarray                  # [K]
    dt0: float
    max_fitness: float
    targets: np.ndarray | None = None   # [B, n_targets]
    obs_keys: np.ndarray | None = None  # [B, 2] uint32
    params: np.ndarray | None = None    # [B, n_params]
    noise_inline: bool = False          # generate the observation noise in-kernel (same bits as the table; tests)
    ys_true: np.ndarray | None = None   # [B, K, D] (SR)
    n_obs: int = 0
    state_size: int = 0
    max_steps: int = 16 ** 4
    obs_noise: float = 0.0
    threefry_partitionable: bool = True
    solver: int = abi.SOLVER_EULER
    controller: int = abi.CONTROLLER_CONSTANT
    rtol: float = 1e-3          # diffrax.PIDController(rtol, atol, dtmin=None, pcoeff=0, icoeff=1, dcoeff=0, ...)
    atol: float = 1e-6
    dtmin: float = 0.0
    pid: tuple = (0.0, 1.0, 0.0)            # pcoeff, icoeff, dcoeff
    pid_limits: tuple = (0.9, 0.2, 10.0)    # safety, factormin, factormax
    _keep: list = field(default_factory=list, repr=False)

    def __post_init__(self) -> None:
        self.x0 = np.ascontiguousarray(self.x0, dtype=np.float32)
        self.ts = np.ascontiguousarray(self.ts, dtype=np.float32)
        if self.ts.ndim != 1 or self.ts.size < 1 or not np.isfinite(self.ts).all() or (np.diff(self.ts) < 0).any():
            raise ValueError("ts must be a non-empty, finite, non-decreasing vector of save times (diffrax.SaveAt(ts=ts))")
        B, D = self.x0.shape
        if self.targets is None:
            self.targets = np.zeros((B, 0), dtype=np.float32)
        self.targets = np.ascontiguousarray(self.targets, dtype=np.float32).reshape(B, -1)
        if self.obs_keys is not None:
            self.obs_keys = np.ascontiguousarray(self.obs_keys, dtype=np.uint32).reshape(B, 2)
        if self.params is not None:
            self.params = np.ascontiguousarray(self.params, dtype=np.float32).reshape(B, -1)
        if self.ys_true is not None:
            self.ys_true = np.ascontiguousarray(self.ys_true, dtype=np.float32)
        if self.kind == abi.EVAL_SR:
            self.n_obs = D
        elif self.n_obs == 0:
            self.n_obs = D

    @property
    def B(self) -> int:
        return self.x0.shape[0]

    @property
    def D(self) -> int:
        return self.x0.shape[1]

    @property
    def K(self) -> int:
        return self.ts.shape[0]

    def fill_desc(self, ptr) -> abi.RolloutDesc:
        """``ptr(array_or_None) -> address`` decides host vs device addresses."""
        d = abi.RolloutDesc()
        d.kind, d.env_id = self.kind, self.env_id
        d.B, d.K, d.D = self.B, self.K, self.D
        d.n_obs = self.n_obs
        d.n_targets = self.targets.shape[1]
        d.n_params = 0 if self.params is None else self.params.shape[1]
        d.state_size = self.state_size
        d.max_steps = int(self.max_steps)
        d.dt0 = float(self.dt0)
        d.max_fitness = float(self.max_fitness)
        d.obs_noise = float(self.obs_noise)
        d.threefry_partitionable = int(bool(self.threefry_partitionable))
        d.solver, d.controller = int(self.solver), int(self.controller)
        d.rtol, d.atol, d.dtmin = float(self.rtol), float(self.atol), float(self.dtmin)
        d.pcoeff, d.icoeff, d.dcoeff = (float(v) for v in self.pid)
        d.pid_safety, d.pid_factormin, d.pid_factormax = (float(v) for v in self.pid_limits)
        d.x0 = ptr(self.x0)
        d.ts = ptr(self.ts)
        d.targets = ptr(self.targets if self.targets.size else None)
        d.obs_keys = ptr(self.obs_keys)
        d.params = ptr(self.params)
        d.params_host = abi.np_ptr(self.params)  # lets the library see trajectory-uniform parameters
        d.flags = abi.DESC_NOISE_INLINE if self.noise_inline else 0
        d.ys_true = ptr(self.ys_true)
        return d

    def to_desc_host(self):
        """Descriptor with HOST pointers (numpy memory).  Returns (desc, keepalive)."""
        return self.fill_desc(abi.np_ptr), self
