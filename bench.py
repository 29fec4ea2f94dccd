#!/usr/bin/env python
"""bench.py -- throughput of the population fitness-evaluation hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c5|c3|c2] [--scaling strong|weak]

One "step" = one pass of the hot path over one synthetic population: encode -> lower -> fused
rollout (-> NCCL all-gather of the fitness shards for N > 1), i.e. what
GeneticProgramming.evaluate_population does once per generation (genetic_programming.py:403-433).

Default workload: BASELINE.json configs[4], "DynamicPolicy rollout sweep": population 65 536 x 32
initial conditions x 2 000 Euler steps, 3 trees (2 latent + 1 readout) of depth <= 6 (max_nodes 63),
Acrobot with observation noise 0.1 as in examples/DynamicPolicy.ipynb.  Under torchrun the named
population is SPLIT over the ranks like the reference's shard_map over the population axis
(genetic_programming.py:255-262, :412-415) -- strong scaling, `scaling: "strong"` -- and the fitness
shards are all-gathered; the same line carries the weak-scaling figure (every rank its own
65 536 individuals) as `weak`.  Also in the line: the pointwise-regression scale-up (configs[3]) as
`secondary` (sharded the same way), the roofline of the dominant kernel, the CPU baseline, the
end-to-end number, and harder / other populations under `extra_workloads`.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from multitreegp_b200 import abi  # noqa: E402
from multitreegp_b200.environments.control_environments import Acrobot  # noqa: E402
from multitreegp_b200.node_library import NodeLibrary  # noqa: E402
from multitreegp_b200.rollout_data import RolloutData  # noqa: E402
from bench_support.synthetic import pointwise_data, random_keys, sample_population  # noqa: E402

OPS_CONTROL = [("+", 2, 0.5), ("-", 2, 0.1), ("*", 2, 0.5), ("sin", 1, 0.1), ("cos", 1, 0.1)]
OPS_SR = [("+", 2, 0.5), ("-", 2, 0.1), ("*", 2, 0.5), ("/", 2, 0.1)]

WORKLOADS = {
    # name: (description, kind, P, B, depth, N, T_end, dt_save, dt0, state_size)
    "c5": ("DynamicPolicy rollout sweep: pop 65536 x 32 ICs x 2000 steps, depth<=6 (BASELINE configs[4])",
           abi.EVAL_DYNAMIC, 65536, 32, 6, 63, 50.2, 0.2, 0.025, 2),
    "c3": ("DynamicPolicy.ipynb: pop 4096 x 32 ICs (BASELINE configs[2])",
           abi.EVAL_DYNAMIC, 4096, 32, 4, 30, 50.2, 0.2, 0.05, 2),
    "c2": ("StaticPolicy.ipynb: pop 1000 x 16 ICs (BASELINE configs[1])",
           abi.EVAL_STATIC, 1000, 16, 4, 30, 50.0, 0.2, 0.05, 0),
}
C4 = dict(P=100_000, n_points=100_000, n_vars=2, depth=8, N=255)

FP32_LANES_PER_SM, N_SM = 128, 148


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except OSError:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


NCU_FILE = os.path.join(ROOT, "profiles", "r02_traffic.json")  # written by tools/ncu_traffic.py from the ncu capture
                                                                  # of THIS build (tools/gpu_final.sh); see its "commit"


def ncu_field(kernel, field):
    try:
        with open(NCU_FILE) as f:
            return json.load(f)[kernel].get(field)
    except (OSError, KeyError):
        return None


def ncu_traffic(kernel, P):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the ncu --set full capture
    recorded in profiles/r02_traffic.json (bytes per individual x P); None when no capture is on file."""
    try:
        with open(NCU_FILE) as f:
            t = json.load(f)[kernel]
        return t["dram_bytes_per_individual"] * P
    except (OSError, KeyError):
        return None


# ---- workloads ---------------------------------------------------------------------------------------
def control_library(env, kind, state_size):
    obs = [f"y{i + 1}" for i in range(env.n_obs)]
    if kind == abi.EVAL_STATIC:
        return NodeLibrary(OPS_CONTROL, [obs], [env.n_control])
    lat = [f"a{i + 1}" for i in range(state_size)]
    return NodeLibrary(OPS_CONTROL, [obs + lat + ["u"], lat], [state_size, env.n_control])  # DynamicPolicy.ipynb:100-102


def make_rollout_workload(name, seed, P_override=None, obs_noise=0.1, trees="fresh"):
    """trees: "fresh" = the reference's initial sampling (operator probability 0.7 ** depth,
    initialization.py:35); "evolved" = bloated trees as after many generations (0.95 ** depth: ~33 live rows per
    tree at depth 6); "full" = every one of the 2^depth - 1 rows live (binary operators only)."""
    desc, kind, P, B, depth, N, T_end, dt_save, dt0, S = WORKLOADS[name]
    P = P_override or P
    rng = np.random.default_rng(seed)
    env = Acrobot(0.05, obs_noise)
    lib = control_library(env, kind, S)
    pop = sample_population(rng, P, lib, depth, N, full=(trees == "full"), op_growth=0.95 if trees == "evolved" else 0.7)
    x0, targets = env.sample_init_states(B, rng)
    ts = np.arange(0, T_end, dt_save, dtype=np.float32)
    params = np.stack(env.sample_params(B, "Constant", ts, rng), axis=1)
    data = RolloutData(kind=kind, env_id=env.env_id, x0=x0, ts=ts, dt0=dt0, max_fitness=1e4, targets=targets,
                       obs_keys=random_keys(rng, B), params=params, n_obs=env.n_obs, state_size=S,
                       obs_noise=obs_noise)
    return dict(name=name, desc=desc, lib=lib, pop=pop, data=data, env=env, kind=kind, B=B, N=N, S=S,
                n_sched=len(_schedule(ts, dt0)))


def _schedule(ts, dt0):
    """Number of Euler steps of the diffrax loop in float32 (same arithmetic as the kernel)."""
    f = np.float32
    t1, tprev, dt0 = f(ts[-1]), f(ts[0]), f(dt0)
    tnext = min(f(tprev + dt0), t1)
    out = []
    while tprev < t1:
        out.append(tprev)
        tprev = min(tnext, t1)
        nn = f(tnext + dt0)
        tnext = t1 if nn > f(t1 - f(1e-6)) else nn
    return out


def algorithmic_ops_per_step(w):
    """algorithmic flops of ONE rollout-step, every add/sub/mul/div/trig = 1:
    F_env (Acrobot drift 75 as written) + tree operators + obs C@x n_obs(2D-1) [+ n_obs(2 n_obs-1) + n_obs
    for the noise term] + Euler 2(D+S) + the post-pass share K/S of (obs + readout operators + cost).
    The threefry/erfinv work of the observation noise is NOT counted: this implementation computes the
    noise once per evaluation (it does not depend on the individual), not once per individual."""
    pop, lib, data = w["pop"], w["lib"], w["data"]
    arity = lib.slots[np.clip(pop[..., 0].astype(np.int64), 0, lib.n_ids - 1)]
    tree_ops = float((arity > 0).sum()) / pop.shape[0]           # operator rows per individual (all trees)
    readout_ops = float((arity[:, w["S"]:] > 0).sum()) / pop.shape[0]
    D, n_obs, S = data.D, data.n_obs, w["S"]
    # C @ x as written (control_environment_base.py:46) + one add per observed component for the noise term.  The
    # reference also forms normal(...) @ W per call (n_obs * (2 n_obs - 1) flops); the kernels read that product from
    # a table computed once per evaluation, so it is NOT counted as work of a rollout-step.
    obs = n_obs * (2 * D - 1) + (n_obs if data.obs_noise != 0.0 else 0)
    post = (data.K / max(w["n_sched"], 1)) * (obs + readout_ops + 8.0)     # reached test (6) + u R u (2)
    flops = 75.0 + tree_ops + obs + 2.0 * (D + S) + post
    return flops, 0.0, tree_ops


# ---- clocks ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---- CPU arms ---------------------------------------------------------------------------------------------
def host_threads():
    """All host cores this process may use (torchrun exports OMP_NUM_THREADS=1; the CPU arm must not obey it)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_sample_size(w, target_s, n_threads=0):
    """How many individuals the oracle evaluates in about target_s seconds on this host."""
    from oracle import pyoracle as po
    po.build()
    n_threads = n_threads or host_threads()
    n_probe = max(32, 8 * n_threads)  # several individuals per thread, or the probe only measures imbalance
    probe = w["pop"][:n_probe]
    t = time.perf_counter(); po.rollout_population(probe, w["lib"].table(), w["data"], n_threads)
    dt = time.perf_counter() - t
    n = int(min(w["pop"].shape[0], max(32, n_probe * target_s / max(dt, 1e-4))))
    return max(32, n - n % 8)


def cpu_time_rollout(w, target_s, n_threads=0, n=None):
    """Times the oracle (CPU restatement, all host threads) on a bounded sample of the workload."""
    from oracle import pyoracle as po
    n_threads = n_threads or host_threads()
    n = n or cpu_sample_size(w, target_s, n_threads)
    sample = w["pop"][:n]
    t = time.perf_counter(); out = po.rollout_population(sample, w["lib"].table(), w["data"], n_threads)
    dt = time.perf_counter() - t
    return out["steps"] / dt, n, out["steps"], dt, n_threads


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path, on all host cores, on a bounded sample
    of the same workload.  When jax + diffrax and the reference sources are present (tools/reference_runner.py:
    $MTGP_REFERENCE, baseline/_ref, /root/reference) this is the REAL reference -- GeneticProgramming.jit_eval
    on JAX-CPU (kind "reference").  They are not installable in this image (no wheels, no network), so here it
    is the oracle port (oracle/mtgp_oracle.c: a C restatement of the reference path, OpenMP over individuals;
    kind "port") -- NOT JAX; say so next to every ratio."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    P_named = WORKLOADS[args.workload][2]
    w = make_rollout_workload(args.workload, seed=1000, P_override=4096, obs_noise=args.obs_noise)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import reference_runner as rr
    real, why = rr.available()
    times, steps_total, kind, cores = [], 0, "port", host_threads()
    if real:
        kind = "reference"
        env = w["env"]
        obs = [f"y{i + 1}" for i in range(env.n_obs)]
        lat = [f"a{i + 1}" for i in range(w["S"])]
        lib_args = ((OPS_CONTROL, [obs], [env.n_control]) if w["kind"] == abi.EVAL_STATIC
                    else (OPS_CONTROL, [obs + lat + ["u"], lat], [w["S"], env.n_control]))
        n = 64
        case = rr.ReferenceCase(w["lib"], *lib_args, w["pop"][:n], w["data"], "acrobot")
        case.fitness()  # compile
        for i in range(args.warmup + args.steps):
            t = time.perf_counter(); case.fitness(); dt = time.perf_counter() - t
            if i >= args.warmup:
                times.append(dt); steps_total += n * w["B"] * w["n_sched"]  # nominal: the reference's batched loop runs every lane to the end
        note = f"the reference itself on JAX-CPU ({rr.versions()})"
    else:
        n = cpu_sample_size(w, args.ref_step_seconds)
        for i in range(args.warmup + args.steps):
            rate, n, steps, dt, cores = cpu_time_rollout(w, args.ref_step_seconds, n=n)
            if i >= args.warmup:
                times.append(dt); steps_total += steps
        note = f"CPU oracle port of the reference path, NOT JAX ({why})"
    value = steps_total / sum(times)
    sample = (f"{n} of {P_named} individuals ({100.0 * n / P_named:.2f} % of the population) x {w['B']} trajectories x "
              f"{w['n_sched']} steps per step; the rate is extrapolated to the population")
    line = {"impl": "reference", "metric": "rollout_steps_per_s", "value": value, "unit": "rollout-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["desc"], "population": P_named, "population_per_gpu": P_named,
                       "trees": int(w["pop"].shape[1]), "tree_sampling": "fresh", "max_nodes": w["N"],
                       "trajectories": w["B"], "euler_steps": w["n_sched"], "obs_noise": args.obs_noise, "env": "Acrobot",
                       "sample_individuals": n, "sample_fraction": n / P_named, "note": note},
            "cpu_baseline": {"value": value, "unit": "rollout-steps/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "rollout-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---- GPU arm --------------------------------------------------------------------------------------------------
def timed_rollout(eng, dev, table, pop_d, data_d, world, K, W, fitness_all=None):
    """W warm-up + K timed evaluations (+ all-gather), bracketed by barrier + synchronize; device time, max over
    ranks.  Returns (elapsed_ms max over ranks, executed rollout-steps summed over ranks, last local fitness)."""
    import torch
    import torch.distributed as dist

    def step():
        out = eng.rollout_fitness(pop_d, table, data_d, count_steps=True)
        if world > 1:
            dist.all_gather_into_tensor(fitness_all, out["fitness"])
        return out

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(W):
        step()
    sync_all()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    ev0.record()
    outs = [step() for _ in range(K)]
    ev1.record()
    sync_all()
    elapsed_ms = ev0.elapsed_time(ev1)
    steps_done = sum(int(o["steps"].item()) for o in outs)
    t = torch.tensor([elapsed_ms, float(steps_done)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        elapsed_ms, steps_done = float(tmax[0]), float(tsum[1])
    return elapsed_ms, steps_done, outs[-1]["fitness"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS))
    ap.add_argument("--population", type=int, default=None, help="override the named (whole-job) population (debug)")
    ap.add_argument("--obs-noise", type=float, default=0.1)
    ap.add_argument("--trees", default="fresh", choices=["fresh", "evolved", "full"])
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="strong (default): the named population is split over the GPUs like the reference's shard_map; "
                         "weak: every GPU evaluates its own full-size population")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--only-secondary", action="store_true", help="run only the pointwise scale-up (profiling)")
    ap.add_argument("--only-notebooks", action="store_true", help="run only the as-shipped notebook configs (profiling)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--ref-step-seconds", type=float, default=2.0)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from multitreegp_b200.distributed import shard_bounds
    from multitreegp_b200.engine import DeviceData, FitnessEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    K, W = args.steps, max(args.warmup, 3)
    peaks, peak_src = measured_peaks()
    if args.only_notebooks:
        print(json.dumps(bench_notebooks(FitnessEngine(dev), dev)))
        return
    if args.only_secondary:
        print(json.dumps(bench_pointwise(FitnessEngine(dev), dev, K, peaks, peak_src, False, world, rank)))
        return

    P_named = args.population or WORKLOADS[args.workload][2]
    eng = FitnessEngine(dev)
    if args.scaling == "strong":
        # ONE population (same seed on every rank), rank r evaluates the contiguous slice [r P/n, (r+1) P/n)
        w = make_rollout_workload(args.workload, seed=1000, P_override=P_named, obs_noise=args.obs_noise, trees=args.trees)
        try:
            lo, hi = shard_bounds(P_named, world, rank)
        except ValueError as e:
            raise SystemExit(str(e))
        w["pop"] = np.ascontiguousarray(w["pop"][lo:hi])
    else:
        w = make_rollout_workload(args.workload, seed=1000 + rank, P_override=P_named, obs_noise=args.obs_noise,
                                  trees=args.trees)
        lo, hi = rank * P_named, (rank + 1) * P_named
    P = w["pop"].shape[0]
    table = w["lib"].table()
    pop_pinned = torch.from_numpy(w["pop"]).pin_memory()
    pop_d = pop_pinned.to(dev)
    data_d = DeviceData(w["data"], dev)
    fitness_all = torch.empty((world * P,), dtype=torch.float32, device=dev)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    elapsed_ms, steps_done, fit_local = timed_rollout(eng, dev, table, pop_d, data_d, world, K, W, fitness_all)
    clocks = sampler.stop() if rank == 0 else None
    value = steps_done / (elapsed_ms * 1e-3)

    # ---- the gathered vector is what every rank's selection step consumes: check it (outside the timed region) ----
    gather_ok = None
    if world > 1:
        mine = torch.equal(fitness_all[rank * P:(rank + 1) * P].view(torch.int32), fit_local.view(torch.int32))
        csum = fitness_all.view(torch.int32).to(torch.int64).sum().reshape(1)
        cmax, cmin = csum.clone(), csum.clone()
        dist.all_reduce(cmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(cmin, op=dist.ReduceOp.MIN)
        flag = torch.tensor([int(mine and int(cmax) == int(cmin))], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        gather_ok = bool(int(flag))  # every rank holds its own shard in place and all ranks hold the same vector

    # ---- weak scaling as a second field: every rank its own full-size population ----
    weak = None
    if world > 1 and args.scaling == "strong":
        ww = make_rollout_workload(args.workload, seed=2000 + rank, P_override=P_named, obs_noise=args.obs_noise,
                                   trees=args.trees)
        wp = torch.from_numpy(ww["pop"]).to(dev)
        wall = torch.empty((world * P_named,), dtype=torch.float32, device=dev)
        w_ms, w_steps, _ = timed_rollout(eng, dev, table, wp, data_d, world, max(2, min(K, 5)), 3, wall)
        kk = max(2, min(K, 5))
        weak = {"value": w_steps / (w_ms * 1e-3), "unit": "rollout-steps/s", "ms_per_step": w_ms / kk,
                "population_per_gpu": P_named, "steps": kk}
        del wp, wall

    # ---- rollout kernel alone (CUDA events on the launching stream) for the roofline --------------
    n, ns, c, cs, P_, T_, N_ = eng._split(pop_d)
    bc, arrs = eng.encode(pop_d, table)
    import ctypes as C
    desc = data_d.desc()
    ws = eng._workspace(eng.lib.mtgp_workspace_bytes(P, T_, N_, C.byref(desc)))
    fit = torch.empty((P,), dtype=torch.float32, device=dev)
    from multitreegp_b200._lib import check
    roll_ms = []
    for i in range(2 + min(K, 5)):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # mtgp_rollout = lower + rollout; lower is timed separately below and subtracted
        a.record()
        check(eng.lib.mtgp_rollout(bc, P, T_, N_, C.byref(table), C.byref(desc), fit.data_ptr(), None, None,
                                   ws.data_ptr(), ws.numel(), eng._stream()))
        b.record(); torch.cuda.synchronize(dev)
        if i >= 2:
            roll_ms.append(a.elapsed_time(b))
    low_ms, enc_ms = [], []
    n_data = desc.n_obs + desc.n_targets + (w["S"] + 1 if w["kind"] == abi.EVAL_DYNAMIC else 0)
    tmp_base = n_data + 2 * w["S"]
    for i in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); eng.lower(bc, P, T_, N_, table, tmp_base=tmp_base); b.record()
        torch.cuda.synchronize(dev)
        if i >= 1:
            low_ms.append(a.elapsed_time(b))
        a.record(); eng.encode(pop_d, table); b.record()
        torch.cuda.synchronize(dev)
        if i >= 1:
            enc_ms.append(a.elapsed_time(b))
    kernel_ms = float(np.mean(roll_ms)) - float(np.mean(low_ms))
    flops, int_ops, tree_ops = algorithmic_ops_per_step(w)
    per_launch_steps = steps_done / (K * world)
    sm_mhz = float(peaks.get("sm_max_mhz", 1965.0))
    fp32_peak = N_SM * FP32_LANES_PER_SM * 2 * sm_mhz * 1e6 / 1e12
    achieved = per_launch_steps * (flops + int_ops) / (kernel_ms * 1e-3) / 1e12
    roofline = {"bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
                "frac": achieved / fp32_peak, "traffic": ncu_traffic("k_rollout", P), "kernel": "k_rollout<DYNAMIC,ACROBOT,noise>",
                "kernel_ms": kernel_ms, "front_end_ms": {"encode": float(np.mean(enc_ms)), "lower": float(np.mean(low_ms))},
                "ops_per_rollout_step": {"fp": flops, "int_prng": int_ops, "tree_ops": tree_ops},
                "ncu": {"file": os.path.relpath(NCU_FILE, ROOT), "commit": ncu_field("k_rollout", "commit"),
                        "issue_slots_busy_pct": ncu_field("k_rollout", "issue_active_pct"),
                        "executed_fp32_frac_of_fma_peak": ncu_field("k_rollout", "executed_fp32_frac_of_fma_peak"),
                        "warp_instructions_per_warp_step": ncu_field("k_rollout", "warp_instructions_per_warp_step"),
                        "warp_instructions_per_32_lane_steps": ncu_field("k_rollout", "warp_instructions_per_32_lane_steps"),
                        "lane_waste": ncu_field("k_rollout", "lane_waste"),
                        "note": "counters of the ncu --set full capture named by `commit` (not of this run); "
                                "warp_instructions_per_warp_step = smsp__inst_executed.sum / warp-steps actually executed (a warp "
                                "steps while any of its lanes still integrates); ..._per_32_lane_steps = the same sum / (executed "
                                "rollout-steps / 32), which also charges the dead lanes that ride along (lane_waste)"},
                "peak_note": f"non-tensor FP32: 148 SM x 128 lanes x 2 (FMA) x {sm_mhz:.0f} MHz ({peak_src} clock); "
                             "the kernel is bound by instruction issue (interpreter dispatch), not by the FP32 pipe"}

    # ---- end to end: HOST buffers through the evaluator class (H2D + kernels + D2H every step) ------
    from multitreegp_b200.evaluators import dynamic_evaluate, feedforward_evaluate
    if w["kind"] == abi.EVAL_DYNAMIC:
        ev = dynamic_evaluate.Evaluator(w["env"], w["S"], w["data"].dt0, library=w["lib"])
    else:
        ev = feedforward_evaluate.Evaluator(w["env"], w["data"].dt0, library=w["lib"])
    pop_host = pop_pinned.numpy()
    e2e_all = torch.empty((world * P,), dtype=torch.float32, device=dev)

    def e2e_step():
        f = ev.evaluate_population(pop_host, w["data"])
        if world > 1:  # the shards' fitness, gathered for the replicated selection step
            dist.all_gather_into_tensor(e2e_all, torch.from_numpy(f).to(dev))
        return f

    for _ in range(2):
        e2e_step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(K):
        fit_host = e2e_step()
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = steps_done / float(te[0])
    h2d = pop_host.nbytes + sum(a.nbytes for a in (w["data"].x0, w["data"].ts, w["data"].obs_keys, w["data"].params))
    e2e = {"value": e2e_value, "unit": "rollout-steps/s", "h2d_bytes_per_step": int(h2d) * world,
           "d2h_bytes_per_step": int(fit_host.nbytes) * world}

    secondary = None
    if not args.no_secondary:
        secondary = bench_pointwise(eng, dev, max(2, min(K, 5)), peaks, peak_src, (not args.no_cpu) and world == 1, world, rank)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    line = {"metric": "rollout_steps_per_s", "value": value, "unit": "rollout-steps/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["desc"], "population": P * world, "population_per_gpu": P, "trees": int(w["pop"].shape[1]),
                       "tree_sampling": args.trees, "max_nodes": w["N"], "trajectories": w["B"], "euler_steps": w["n_sched"],
                       "obs_noise": args.obs_noise, "env": "Acrobot",
                       "nominal_rollout_steps_per_step": P * w["B"] * w["n_sched"] * world,
                       "executed_rollout_steps_per_step": steps_done / K,
                       "l2": f"inputs larger than L2 at 1 GPU (population {pop_host.nbytes / 1e6:.0f} MB per GPU re-read from HBM "
                             "every step; 126 MB L2)",
                       "step": "encode + lower + fused rollout" + (" + NCCL all-gather" if world > 1 else "")},
            "clocks": clocks, "e2e": e2e, "gpu_launches": (7 if args.obs_noise != 0.0 else 5) * K, "roofline": roofline}
    if gather_ok is not None:
        line["gather_verified"] = gather_ok
    if weak is not None:
        line["weak"] = weak
    if secondary is not None:
        line["secondary"] = secondary

    if not args.no_secondary and world == 1:
        line["notebook_configs"] = bench_notebooks(eng, dev)
        line["extra_workloads"] = bench_extra(eng, dev, args)
    if not args.no_cpu and world == 1:
        rate, n, steps, dt, cores = cpu_time_rollout(w, target_s=12.0)
        line["cpu_baseline"] = {"value": rate, "unit": "rollout-steps/s", "cores": cores, "kind": "port",
                                "sample": f"{n} of {P} individuals x {w['B']} trajectories x {w['n_sched']} steps "
                                          f"({steps} rollout-steps in {dt:.1f} s = {100.0 * n / P:.2f} % of the population; "
                                          "oracle/mtgp_oracle.c with OpenMP, a C port of the reference path, NOT JAX)"}
        # the same port doing the WORK the reference does under vmap o switch: every node function on every one of the
        # max_nodes rows (BASELINE.md 3.3 "reference-faithful"); a quarter of the sample, same values
        from oracle import pyoracle as po
        po.lib().oracle_set_all_branches(1)
        try:
            rate_f, n_f, steps_f, dt_f, _ = cpu_time_rollout(w, target_s=0.0, n=max(32, (n // 4) - (n // 4) % 8))
        finally:
            po.lib().oracle_set_all_branches(0)
        line["cpu_baseline"]["all_branches"] = {
            "value": rate_f, "unit": "rollout-steps/s",
            "sample": f"{n_f} individuals ({steps_f} rollout-steps in {dt_f:.1f} s); every node function evaluated on "
                      "every row, as XLA does for lax.switch under vmap (genetic_programming.py:372)"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def bench_extra(eng, dev, args):
    """Populations other than the freshly sampled C5 one: the same C5 shapes with evolved-size
    and full trees, the SymbolicRegression notebook shapes (C1), each as ms per evaluation of the whole population."""
    import torch
    from multitreegp_b200.engine import DeviceData
    out = {}

    def time_eval(pop, table, data_d):
        pop_d = torch.from_numpy(pop).to(dev)
        for _ in range(3):
            eng.rollout_fitness(pop_d, table, data_d, count_steps=True)
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(2):
            o = eng.rollout_fitness(pop_d, table, data_d, count_steps=True)
        b.record(); torch.cuda.synchronize(dev)
        return a.elapsed_time(b) / 2, int(o["steps"].item())

    for trees in ("evolved", "full"):
        w = make_rollout_workload("c5", seed=1000, obs_noise=args.obs_noise, trees=trees)
        ms, steps = time_eval(w["pop"], w["lib"].table(), DeviceData(w["data"], dev))
        out[f"c5_{trees}_trees"] = {"workload": w["desc"], "tree_sampling": trees,
                                    "mean_live_rows_per_tree": float((w["pop"][..., 0] != 0).sum(-1).mean()),
                                    "ms_per_step": ms, "rollout_steps_per_s": steps / (ms * 1e-3)}
    # C1: SymbolicRegression.ipynb:128-144 shapes -- 10 x 100 individuals, 2 trees, max_nodes 30, 16 Van der Pol
    # trajectories saved on arange(0, 20, 0.2); the evaluator's default Euler(dt0 = 0.01) and the notebook's Dopri5 + PID
    from bench_support.synthetic import van_der_pol_data
    rng = np.random.default_rng(11)
    lib = NodeLibrary(OPS_SR, [["x0", "x1"]], [2])
    pop = sample_population(rng, 1000, lib, 4, 30)
    x0, ts, ys, _ = van_der_pol_data(rng, 16, 20.0, 0.2)
    for solver in ("euler", "dopri5_pid"):
        d = RolloutData(kind=abi.EVAL_SR, env_id=abi.ENV_NONE, x0=x0, ts=ts, dt0=0.01, max_fitness=1e5, ys_true=ys)
        if solver == "dopri5_pid":  # SymbolicRegression.ipynb:136
            d.solver, d.controller, d.rtol, d.atol, d.dtmin, d.max_steps = (abi.SOLVER_DOPRI5, abi.CONTROLLER_PID,
                                                                              1e-6, 1e-6, 0.001, 500)
        ms, steps = time_eval(pop, lib.table(), DeviceData(d, dev))
        out[f"c1_{solver}"] = {"workload": "SymbolicRegression.ipynb: pop 10 x 100, 2 trees, 16 trajectories, K = 100 "
                                           "(BASELINE configs[0] shapes)", "solver": solver, "ms_per_step": ms,
                               "solver_steps_per_s": steps / (ms * 1e-3)}
    return out


def bench_notebooks(eng, dev):
    """BASELINE configs[1] and [2] (StaticPolicy / DynamicPolicy notebooks, 1 000 x 16 and 4 096 x 32): ms per
    evaluation with (a) the fixed-step Euler stepper the configs name (the evaluators' class default) and
    (b) the solver settings the notebooks actually pass, Dopri5 + PIDController(1e-4, 1e-4, dtmin=1e-3),
    max_steps 1000.  Small, latency-bound workloads -- reported for completeness, not as the headline."""
    import torch
    from multitreegp_b200.engine import DeviceData
    out = {}
    for name in ("c2", "c3"):
        for solver in ("euler", "dopri5_pid"):
            w = make_rollout_workload(name, seed=5)
            d = w["data"]
            if solver == "dopri5_pid":
                d.solver, d.controller, d.rtol, d.atol, d.dtmin, d.max_steps = (abi.SOLVER_DOPRI5, abi.CONTROLLER_PID,
                                                                                  1e-4, 1e-4, 0.001, 1000)
            pop_d, data_d, table = torch.from_numpy(w["pop"]).to(dev), DeviceData(d, dev), w["lib"].table()
            for _ in range(3):
                eng.rollout_fitness(pop_d, table, data_d, count_steps=True)
            torch.cuda.synchronize(dev)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            o = eng.rollout_fitness(pop_d, table, data_d, count_steps=True)
            b.record(); torch.cuda.synchronize(dev)
            ms = a.elapsed_time(b)
            out[f"{name}_{solver}"] = {"workload": w["desc"], "solver": solver, "ms_per_evaluation": ms,
                                       "solver_steps_per_s": int(o["steps"].item()) / (ms * 1e-3)}
    return out


def bench_pointwise(eng, dev, K, peaks, peak_src, with_cpu, world=1, rank=0):
    """BASELINE configs[3]: pointwise regression, population 100k x 100k points (n_vars 2, and the wide variant
    n_vars 8 at one GPU).  The population is sharded over the ranks like the rollout; value = whole job."""
    import torch
    import torch.distributed as dist
    from multitreegp_b200.distributed import shard_bounds

    def run(n_vars):
        rng = np.random.default_rng(7)
        lib = NodeLibrary(OPS_SR, [[f"x{i}" for i in range(n_vars)]], [1])   # SymbolicRegression.ipynb:138-141
        pop = sample_population(rng, C4["P"], lib, C4["depth"], C4["N"])
        X, Y = pointwise_data(rng, C4["n_points"], n_vars)
        lo, hi = shard_bounds(C4["P"], world, rank)
        live_total = float((pop[..., 0] != 0).sum())
        ops_total = float((lib.slots[np.clip(pop[..., 0].astype(np.int64), 0, lib.n_ids - 1)] > 0).sum())
        pop_d, Xd, Yd = (torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (pop[lo:hi], X, Y))
        table = lib.table()
        gathered = torch.empty((C4["P"],), dtype=torch.float32, device=dev)

        def step():
            f = eng.pointwise_fitness(pop_d, table, Xd, Yd)
            if world > 1:
                dist.all_gather_into_tensor(gathered, f)

        for _ in range(3):
            step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(K):
            step()
        b.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        t = torch.tensor([a.elapsed_time(b) / K], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), live_total, ops_total, lib, pop, X, Y, table

    ms, live, ops, lib, pop, X, Y, table = run(C4["n_vars"])
    node_evals = live * C4["n_points"]
    logical_bytes = C4["P"] * C4["n_points"] * (C4["n_vars"] + 1) * 4.0
    hbm = float(peaks.get("hbm_gbs", 6650.0)) * world
    sm_mhz = float(peaks.get("sm_max_mhz", 1965.0))
    fp32_peak = N_SM * FP32_LANES_PER_SM * 2 * sm_mhz * 1e6 / 1e12 * world
    flops = (ops + 3.0 * C4["P"]) * C4["n_points"]  # one per operator node per point + (pred - y)^2 and its sum
    out = {"metric": "tree_node_evals_per_s", "value": node_evals / (ms * 1e-3), "unit": "live node-evals/s",
           "n_gpus": world, "scaling": "strong",
           "padded_node_evals_per_s": C4["P"] * C4["N"] * C4["n_points"] / (ms * 1e-3), "ms_per_step": ms,
           "config": {"workload": "symbolic regression scale-up: pop 100k x 100k points, depth<=8 (BASELINE configs[3])",
                      "n_vars": C4["n_vars"], "mean_live_nodes": live / C4["P"], "population_per_gpu": C4["P"] // world},
           "roofline": {"bound": "issue", "logical_GBps": logical_bytes / (ms * 1e-3) / 1e9, "hbm_peak_GBps": hbm,
                        "logical_over_hbm_peak": logical_bytes / (ms * 1e-3) / 1e9 / hbm,
                        "algorithmic_TFLOPs": flops / (ms * 1e-3) / 1e12, "fp32_frac": flops / (ms * 1e-3) / 1e12 / fp32_peak,
                        "dram_bytes_per_launch_ncu": ncu_field("k_pointwise", "dram_bytes_per_launch"),
                        "issue_slots_busy_pct_ncu": ncu_field("k_pointwise", "issue_active_pct"),
                        "note": "the 1.2 MB data set is L2-resident: DRAM traffic is the population and its programs, and the "
                                f"kernel is bound by instruction issue / shared memory, not by HBM.  logical_GBps = P x n_points x "
                                f"(n_vars + 1) x 4 B / t is the streaming model of SURVEY 8(d), quoted against the {peak_src} copy "
                                "peak only as that model asks; it is NOT an HBM roofline fraction"}}
    if world == 1:
        # end to end through the evaluator's host call: pinned host population (408 MB), X, Y -> fitness in host memory
        import torch as _t
        pop_h = _t.from_numpy(np.ascontiguousarray(pop)).pin_memory()
        for _ in range(2):
            eng.pointwise_fitness_host(pop_h.numpy(), table, X, Y)
        t0 = time.perf_counter()
        for _ in range(K):
            eng.pointwise_fitness_host(pop_h.numpy(), table, X, Y)
        ms_e = (time.perf_counter() - t0) * 1e3 / K
        out["e2e"] = {"value": node_evals / (ms_e * 1e-3), "unit": "live node-evals/s", "ms_per_step": ms_e,
                      "h2d_bytes_per_step": int(pop.nbytes + X.nbytes + Y.nbytes), "d2h_bytes_per_step": int(4 * C4["P"]),
                      "note": "mtgp_population_fitness_pointwise_host: upload in eight slices, evaluated on two streams as they land"}
        del pop_h
        ms8, live8, *_ = run(8)
        out["wide_variant"] = {"n_vars": 8, "ms_per_step": ms8, "value": live8 * C4["n_points"] / (ms8 * 1e-3),
                               "unit": "live node-evals/s"}
    if with_cpu and rank == 0:
        from oracle import pyoracle as po
        n, npts = 256, 4096
        t = time.perf_counter(); po.pointwise_population(pop[:n], table, X[:npts], Y[:npts], n_threads=host_threads())
        dt = time.perf_counter() - t
        live_s = float((pop[:n, ..., 0] != 0).sum())
        out["cpu_baseline"] = {"value": live_s * npts / dt, "unit": "live node-evals/s", "cores": host_threads(),
                               "kind": "port", "sample": f"{n} individuals x {npts} points ({dt:.1f} s)"}
    return out


if __name__ == "__main__":
    main()
