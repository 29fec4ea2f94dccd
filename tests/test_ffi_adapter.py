"""The XLA-FFI adapter (multitreegp_b200/csrc/mtgp_xla_ffi.cc), built against the stand-in header
tests/fake_xla_ffi/xla/ffi/api/ffi.h (jaxlib's header is not installable in this image).

CPU: the adapter compiles, exports both handler symbols, and the attribute list multitreegp_b200/jax_ffi.py
sends is the list the handler binds (names and order).
GPU: the handler bodies run with buffers shaped as `jax.ffi.ffi_call(..., vmap_method="expand_dims")` under
vmap(shard_map) delivers them -- coefficients [P,T,N,1], nodes [P,T,N,3], un-batched data with a leading 1,
zero-size targets f32[1,B,0] -- and give the bits of the ctypes path.  Reference contract: SURVEY.md 8(b),
genetic_programming.py:249-269."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from multitreegp_b200 import abi
from tests import util

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(ROOT, "multitreegp_b200", "csrc", "mtgp_xla_ffi.cc")
SO = os.path.join(HERE, "hostlib", "libmtgp_ffi_fake.so")


class Buf(C.Structure):
    _fields_ = [("data", C.c_void_p), ("dims", C.POINTER(C.c_int64)), ("rank", C.c_int32)]


@pytest.fixture(scope="module")
def ffi_lib():
    deps = [SRC, os.path.join(HERE, "fake_xla_ffi", "xla", "ffi", "api", "ffi.h"),
            os.path.join(ROOT, "multitreegp_b200", "libmtgp_b200.so")]
    if not os.path.exists(SO) or any(os.path.getmtime(p) > os.path.getmtime(SO) for p in deps):
        subprocess.check_call(["nvcc", "-x", "cu", "-std=c++17", "-O2", "-Xcompiler", "-fPIC", "-Xcompiler", "-w", "-shared",
                               "-I", os.path.join(HERE, "fake_xla_ffi"), SRC, "-o", SO,
                               "-L", os.path.join(ROOT, "multitreegp_b200"), "-lmtgp_b200",
                               "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../../multitreegp_b200"])
    return C.CDLL(SO)


def test_adapter_compiles_and_exports_the_handlers(ffi_lib):
    for sym in ("mtgp_rollout_ffi", "mtgp_pointwise_ffi", "mtgp_rollout_ffi_test", "mtgp_pointwise_ffi_test"):
        assert hasattr(ffi_lib, sym)


def test_python_side_sends_the_attributes_the_handler_binds():
    """jax_ffi.rollout_attributes (what JaxEvaluator passes to ffi_call) == the .Attr<...>() list of mtgp_rollout_ffi."""
    from multitreegp_b200 import jax_ffi, solvers
    from multitreegp_b200.evaluators import dynamic_evaluate
    src = open(SRC).read()
    body = src[src.index("mtgp_rollout_ffi, RolloutImpl"):src.index("// Pointwise evaluator")]
    bound = re.findall(r'\.Attr<[^>]*>+\("(\w+)"\)', body)
    env = util.ENVS["acrobot"](0.05, 0.1)
    lib = util.control_library(env, abi.EVAL_DYNAMIC, 2)
    ev = dynamic_evaluate.Evaluator(env, 2, 0.05, solver=solvers.Dopri5(),
                                    stepsize_controller=solvers.PIDController(1e-4, 1e-5, pcoeff=0.1, icoeff=0.3, dcoeff=0.05,
                                                                              dtmin=0.001, safety=0.8, factormin=0.3,
                                                                              factormax=5.0), max_steps=1000, library=lib)
    attrs = jax_ffi.rollout_attributes(ev, True)
    assert list(attrs) == bound
    assert attrs["solver"] == abi.SOLVER_DOPRI5 and attrs["controller"] == abi.CONTROLLER_PID
    assert (float(attrs["pcoeff"]), float(attrs["icoeff"]), float(attrs["dcoeff"])) == pytest.approx((0.1, 0.3, 0.05))
    assert (float(attrs["pid_safety"]), float(attrs["pid_factormin"]), float(attrs["pid_factormax"])) == pytest.approx((0.8, 0.3, 5.0))


def _bufs(arrs):
    keep, out = [], (Buf * len(arrs))()
    for i, (ptr, dims) in enumerate(arrs):
        d = (C.c_int64 * max(1, len(dims)))(*dims)
        keep.append(d)
        out[i].data, out[i].dims, out[i].rank = ptr, C.cast(d, C.POINTER(C.c_int64)), len(dims)
    return out, keep


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["dynamic_acrobot", "static_harmonic", "sr", "dynamic_acrobot_dopri5_pid"])
def test_rollout_handler_matches_the_ctypes_path(ffi_lib, case):
    import torch
    from multitreegp_b200.engine import DeviceData, FitnessEngine
    rng = np.random.default_rng(5)
    if case == "sr":
        lib, pop, data = util.sr_case(rng, P=40, B=6, T_end=3.0)
    else:
        env, kind = ("harmonic", abi.EVAL_STATIC) if case == "static_harmonic" else ("acrobot", abi.EVAL_DYNAMIC)
        lib, pop, data = util.control_case(env, kind, rng, P=40, B=6, T_end=4.0, obs_noise=0.1)
        if case.endswith("pid"):
            util.with_solver(data, **util.DOPRI_PID)
            data.max_steps = 1000
            data.pid = (0.1, 0.3, 0.05)
            data.pid_limits = (0.8, 0.3, 5.0)
    eng = FitnessEngine()
    dev = eng.device
    want = eng.rollout_fitness(torch.from_numpy(pop).to(dev), lib.table(), DeviceData(data, dev))["fitness"].cpu().numpy()

    P, T, N, _ = pop.shape
    t = lambda a, dt=torch.float32: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    nodes, coef = t(pop[..., :3]), t(pop[..., 3:])                      # the split arrays of shard_eval (:261)
    B, D, K = data.x0.shape[0], data.x0.shape[1], data.ts.shape[0]
    x0, ts = t(data.x0), t(data.ts)
    targets = t(data.targets) if data.targets is not None and data.targets.size else None
    n_t = 0 if targets is None else data.targets.shape[1]
    keys = t(data.obs_keys.view(np.int32)) if data.obs_keys is not None else None
    params = t(data.params) if data.params is not None else None
    n_p = 0 if params is None else data.params.shape[1]
    ys = t(data.ys_true) if data.ys_true is not None else None
    fit = torch.empty(P, dtype=torch.float32, device=dev)
    ptr = lambda x: 0 if x is None else x.data_ptr()  # noqa: E731
    bufs, keep = _bufs([(ptr(coef), [P, T, N, 1]), (ptr(nodes), [P, T, N, 3]), (ptr(x0), [1, B, D]), (ptr(ts), [1, K]),
                        (ptr(targets), [1, B, n_t]), (ptr(keys), [1, B, 2] if keys is not None else [1, B, 0]),
                        (ptr(params), [1, B, n_p]), (ptr(ys), [1, B, K, D] if ys is not None else [0]), (ptr(fit), [P])])
    tab = lib.table()
    opc = (C.c_int32 * tab.n_ids)(*[int(tab.opcode[i]) for i in range(tab.n_ids)])
    arg = (C.c_int32 * tab.n_ids)(*[int(tab.arg[i]) for i in range(tab.n_ids)])
    iattrs = (C.c_int32 * 8)(data.kind, data.env_id, data.n_obs if data.kind != abi.EVAL_SR else D, data.state_size,
                             data.max_steps, int(data.threefry_partitionable), data.solver, data.controller)
    pid, lim = tuple(data.pid), tuple(data.pid_limits)
    fattrs = (C.c_float * 12)(data.dt0, data.max_fitness, data.obs_noise, data.rtol, data.atol, data.dtmin,
                              pid[0], pid[1], pid[2], lim[0], lim[1], lim[2])
    err = C.create_string_buffer(256)
    rc = ffi_lib.mtgp_rollout_ffi_test(C.c_void_p(torch.cuda.current_stream(dev).cuda_stream), bufs, opc, arg,
                                       tab.n_ids, iattrs, fattrs, err, 256)
    assert rc == 0, err.value
    got = fit.cpu().numpy()
    assert util.same_or_both_nan(got, want).all()
    del keep


@pytest.mark.gpu
def test_handler_reports_invalid_arguments(ffi_lib):
    import torch
    dev = torch.device("cuda")
    z = torch.zeros(16, device=dev)
    bufs, keep = _bufs([(z.data_ptr(), [2, 1, 2, 1]), (z.data_ptr(), [2, 1, 2, 2]), (0, [0]), (0, [0]), (0, [0]), (0, [0]),
                        (0, [0]), (0, [0]), (z.data_ptr(), [2])])
    err = C.create_string_buffer(256)
    one = (C.c_int32 * 1)(0)
    rc = ffi_lib.mtgp_rollout_ffi_test(None, bufs, one, one, 1, (C.c_int32 * 8)(), (C.c_float * 12)(), err, 256)
    assert rc == 3 and b"nodes must be" in err.value   # ffi::Error::InvalidArgument
    del keep


@pytest.mark.gpu
def test_pointwise_handler_matches_the_ctypes_path(ffi_lib):
    import torch
    from bench_support.synthetic import pointwise_data, sample_population
    from multitreegp_b200.engine import FitnessEngine
    from multitreegp_b200.node_library import NodeLibrary
    rng = np.random.default_rng(9)
    lib = NodeLibrary(util.OPS_SR, [["x0", "x1"]], [1])
    pop = sample_population(rng, 64, lib, 5, 31)
    X, Y = pointwise_data(rng, 1000, 2)
    eng = FitnessEngine()
    dev = eng.device
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    want = eng.pointwise_fitness(t(pop), lib.table(), t(X), t(Y)).cpu().numpy()
    P, T, N, _ = pop.shape
    nodes, coef, Xd, Yd = t(pop[..., :3]), t(pop[..., 3:]), t(X), t(Y)
    fit = torch.empty(P, dtype=torch.float32, device=dev)
    bufs, keep = _bufs([(coef.data_ptr(), [P, T, N, 1]), (nodes.data_ptr(), [P, T, N, 3]), (Xd.data_ptr(), [1, 1000, 2]),
                        (Yd.data_ptr(), [1, 1000, T]), (fit.data_ptr(), [P])])
    tab = lib.table()
    opc = (C.c_int32 * tab.n_ids)(*[int(tab.opcode[i]) for i in range(tab.n_ids)])
    arg = (C.c_int32 * tab.n_ids)(*[int(tab.arg[i]) for i in range(tab.n_ids)])
    err = C.create_string_buffer(256)
    rc = ffi_lib.mtgp_pointwise_ffi_test(C.c_void_p(torch.cuda.current_stream(dev).cuda_stream), bufs, opc, arg, tab.n_ids,
                                         C.c_float(1e5), err, 256)
    assert rc == 0, err.value
    assert util.same_or_both_nan(fit.cpu().numpy(), want).all()
    del keep
