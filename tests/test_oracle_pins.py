"""What pins the CPU oracle in the absence of a runnable reference:
known-answer vectors, hand-computed cases, closed forms and an independent NumPy restatement."""
import numpy as np
import pytest

from multitreegp_b200 import abi
from multitreegp_b200.environments.control_environments import Acrobot, HarmonicOscillator
from multitreegp_b200.node_library import NodeLibrary
from multitreegp_b200.rollout_data import RolloutData
from bench_support.synthetic import sample_population
from tests import np_reference, util

F32 = np.float32


# ---- jax.random ------------------------------------------------------------------------------------
def test_threefry2x32_known_answers(oracle):
    # Random123 kat_vectors, threefry2x32 20 rounds
    assert oracle.threefry2x32(0, 0, 0, 0) == (0x6B200159, 0x99BA4EFE)
    assert oracle.threefry2x32(0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF) == (0x1CB996FC, 0xBB002BE7)
    assert oracle.threefry2x32(0x13198A2E, 0x03707344, 0x243F6A88, 0x85A308D3) == (0xC4923A9C, 0x483DF7A0)


def test_fold_in_and_normal_layouts(oracle):
    key = (0x12345678, 0x9ABCDEF0)
    t = F32(0.35)
    bits_t = int(np.float32(t).view(np.uint32))
    fk = oracle.fold_in(key, bits_t)
    assert fk == oracle.threefry2x32(key[0], key[1], 0, bits_t)

    def to_normal(bits):
        f = (np.uint32((bits >> 9) | 0x3F800000)).view(F32) - F32(1.0)
        lo = np.nextafter(F32(-1), F32(0))
        u = np.maximum(lo, F32(f * F32(F32(1.0) - lo) + lo))
        return F32(np.sqrt(2)) * oracle.math_apply("erfinv", F32([u]))[0]

    for n in (2, 3, 4):
        part = oracle.normal(fk, n, True)
        for i in range(n):
            o = oracle.threefry2x32(fk[0], fk[1], 0, i)
            assert part[i] == to_normal(o[0] ^ o[1])
        orig = oracle.normal(fk, n, False)
        half = (n + 1) // 2
        counts = list(range(n)) + ([0] if n % 2 else [])
        outs = [oracle.threefry2x32(fk[0], fk[1], counts[i], counts[half + i]) for i in range(half)]
        flat = [o[0] for o in outs] + [o[1] for o in outs]
        for i in range(n):
            assert orig[i] == to_normal(flat[i])


def test_normal_is_standard_normal(oracle):
    rng = np.random.default_rng(0)
    v = np.concatenate([oracle.normal(tuple(int(x) for x in rng.integers(0, 2 ** 32, 2)), 4) for _ in range(5000)])
    assert abs(v.mean()) < 0.03 and abs(v.std() - 1) < 0.03 and np.isfinite(v).all()


# ---- tree interpreter ----------------------------------------------------------------------------
def _static_lib():
    return NodeLibrary(util.OPS_CONTROL, [["y1", "y2", "y3", "y4"]], [1])


def test_worked_example_of_the_survey(oracle):
    """y4 + sin(0.9*y4) in StaticPolicy ids (+=2 -=3 *=4 sin=5 cos=6, y1..y4 = 7..10), SURVEY.md App. A."""
    lib = _static_lib()
    assert lib.string_to_node == {"+": 2, "-": 3, "*": 4, "sin": 5, "cos": 6, "y1": 7, "y2": 8, "y3": 9, "y4": 10}
    N = 30
    tree = np.tile(F32([0, -1, -1, 0]), (N, 1))
    tree[29] = [2, 28, 27, 0]
    tree[28] = [10, -1, -1, 0]
    tree[27] = [5, 26, -1, 0]
    tree[26] = [4, 25, 24, 0]
    tree[25] = [1, -1, -1, 0.9]
    tree[24] = [10, -1, -1, 0]
    data = F32([0.1, 0.2, 0.3, 1.7])
    got = oracle.tree_eval(tree, lib.table(), data)
    oracle.set_math(oracle.MATH_LIBM)
    got_libm = oracle.tree_eval(tree, lib.table(), data)
    oracle.set_math(oracle.MATH_CONTRACT)
    want = F32(1.7) + np.sin(F32(F32(0.9) * F32(1.7)))
    assert abs(got - want) <= 2 * np.spacing(F32(want)) and got_libm == F32(want)


def test_interpreter_edge_semantics(oracle):
    lib = NodeLibrary(util.OPS_SR, [["x0", "x1"]], [1])  # + - * / = 2..5, x0 = 6, x1 = 7
    tab, N = lib.table(), 8
    empty = np.tile(F32([0, -1, -1, 0]), (N, 1))
    data = F32([3.0, 0.0])
    assert oracle.tree_eval(empty, tab, data) == 0.0                       # all rows empty -> 0.0
    t = empty.copy(); t[-1] = [1, -1, -1, -2.5]
    assert oracle.tree_eval(t, tab, data) == F32(-2.5)                     # coefficient leaf
    t = empty.copy(); t[-1] = [5, 6, 5, 0]; t[6] = [6, -1, -1, 0]; t[5] = [7, -1, -1, 0]
    assert oracle.tree_eval(t, tab, data) == np.inf                        # x0 / x1 = 3/0, unprotected
    t[6] = [7, -1, -1, 0]
    assert np.isnan(oracle.tree_eval(t, tab, data))                        # 0/0
    t = empty.copy(); t[-1] = [3, 6, 5, 0]; t[6] = [6, -1, -1, 0]; t[5] = [1, -1, -1, 1.0]
    assert oracle.tree_eval(t, tab, data) == F32(2.0)                      # x - y operand order: a is the left operand
    t = empty.copy(); t[-1] = [99, -1, -1, 0]
    assert oracle.tree_eval(t, tab, data) == F32(0.0)                      # id clamps to the last function (x1)
    t = empty.copy(); t[-1] = [1.5, -1, -1, 7.0]
    assert oracle.tree_eval(t, tab, data) == F32(0.0)                      # int(1.5)=1 -> fn[1]=0.0, but f != 1
    # a = -1 on an operator reads the LAST row's current column 3 (its initial coefficient field)
    t = empty.copy(); t[-1] = [2, -1, 6, 4.0]; t[6] = [6, -1, -1, 0]
    assert oracle.tree_eval(t, tab, data) == F32(4.0 + 3.0)


@pytest.mark.parametrize("ops,tol", [(util.OPS_SR, 0), (util.OPS_ALL, 4)])
def test_oracle_vs_numpy_restatement(oracle, ops, tol):
    rng = np.random.default_rng(1)
    lib = NodeLibrary(ops, [["x0", "x1", "x2"]], [1])
    fns = np_reference.node_functions(lib)
    pop = sample_population(rng, 150, lib, 5, 40)
    data = F32(rng.normal(size=(4, 3)) * 2)
    oracle.set_math(oracle.MATH_LIBM if tol else oracle.MATH_CONTRACT)
    try:
        for p in range(pop.shape[0]):
            for d in data:
                a, b = oracle.tree_eval(pop[p, 0], lib.table(), d), np_reference.foriloop(pop[p, 0], d, fns)
                if np.isnan(a) or np.isnan(b):
                    assert np.isnan(a) and np.isnan(b)
                elif tol == 0 or not np.isfinite(b):
                    assert a == b
                else:  # libm vs numpy transcendental functions may differ by an ulp, amplified through the tree
                    assert abs(a - b) <= 1e-4 * max(1.0, abs(b))
    finally:
        oracle.set_math(oracle.MATH_CONTRACT)


# ---- diffrax loop --------------------------------------------------------------------------------
def test_step_schedule(oracle):
    ts = np.arange(0, 50, 0.2, dtype=F32)
    tp, tn = oracle.schedule(ts, 0.05)
    assert tp[0] == ts[0] and tn[-1] == ts[-1] and 990 <= tp.size <= 1000
    assert (tn[:-1] == tp[1:]).all() and (np.diff(tn) > 0).all()
    dt0 = F32(0.05)
    acc = F32(ts[0] + dt0)
    for k in range(tn.size - 1):           # ConstantStepSize: tnext accumulates in float32
        assert tn[k] == acc
        acc = F32(acc + dt0)
    tp2, _ = oracle.schedule(ts, 0.05, max_steps=37)
    assert tp2.size == 37
    tp3, tn3 = oracle.schedule(F32([0, 1, 2]), 0.75)   # dt0 does not divide the horizon: last step clipped
    np.testing.assert_array_equal(tn3, F32([0.75, 1.5, 2.0]))


def _const_policy_pop(lib, N, value):
    tree = np.tile(F32([0, -1, -1, 0]), (N, 1))
    tree[-1] = [1, -1, -1, value]
    return tree[None, None].copy()


def test_harmonic_oscillator_closed_form(oracle):
    """Constant policy u = c: Euler is the linear recurrence x' = (I + hA) x + h b c; times are exact
    binary fractions so every save point coincides with a step."""
    env = HarmonicOscillator(0.0, 0.0)
    lib = util.control_library(env, abi.EVAL_STATIC)
    B, h, c = 5, 0.25, 0.75
    rng = np.random.default_rng(2)
    x0, targets = env.sample_init_states(B, rng)
    omega, zeta = env.sample_params(B, "Different", None, rng)
    ts = np.arange(0, 8, 0.5, dtype=F32)
    data = RolloutData(kind=abi.EVAL_STATIC, env_id=env.env_id, x0=x0, ts=ts, dt0=h, max_fitness=1e4,
                       targets=targets, params=np.stack([omega, zeta], 1), n_obs=2)
    out = oracle.rollout_population(_const_policy_pop(lib, 12, c), lib.table(), data, trajectories=True)
    for b in range(B):
        A = np.array([[0, 1], [-omega[b], -zeta[b]]], dtype=np.float64)
        x = x0[b].astype(np.float64)
        xs = [x.copy()]
        for _ in range(int(round(ts[-1] / h))):
            x = x + h * (A @ x + np.array([0, 1.0]) * c)
            xs.append(x.copy())
        xs = np.array(xs)[::2]
        np.testing.assert_allclose(out["xs"][0, b], xs, rtol=2e-5, atol=2e-5)
        xd = np.array([targets[b, 0], 0.0])
        ud = omega[b] * targets[b, 0]
        cost = sum(0.5 * (s[0] - xd[0]) ** 2 + 0.5 * (c - ud) ** 2 for s in xs)
        np.testing.assert_allclose(out["traj_fitness"][0, b], cost, rtol=2e-5)
        np.testing.assert_allclose(out["us"][0, b, :, 0], c)
    want = np.clip(np.mean(np.where(np.isfinite(out["traj_fitness"][0]), out["traj_fitness"][0], 1e4)), 0, 1e4)
    np.testing.assert_allclose(out["fitness"][0], want, rtol=1e-6)


def test_sr_true_field_matches_independent_euler(oracle):
    """Trees = the Van der Pol field; an independent float32 Euler in NumPy must agree bit for bit."""
    lib = NodeLibrary(util.OPS_SR, [["x0", "x1"]], [2])  # + - * / = 2..5, x0 = 6, x1 = 7
    N = 12
    t0 = np.tile(F32([0, -1, -1, 0]), (N, 1)); t0[-1] = [7, -1, -1, 0]          # dx0 = x1
    # dx1 = ((1 - x0*x0) * x1) - x0, postfix: right subtree first
    rows = [[6, -1, -1, 0],                       # x0                      (right operand of the outer -)
            [7, -1, -1, 0],                       # x1                      (right operand of *)
            [6, -1, -1, 0], [6, -1, -1, 0], [4, 3, 2, 0],   # x0, x0, x0*x0
            [1, -1, -1, 1.0],                     # 1
            [3, 5, 4, 0],                         # 1 - x0*x0
            [4, 6, 1, 0],                         # (1 - x0*x0) * x1
            [3, 7, 0, 0]]                         # ... - x0
    t1 = np.tile(F32([0, -1, -1, 0]), (N, 1))
    off = N - len(rows)
    for i, r in enumerate(rows):
        t1[off + i] = [r[0], r[1] + off if r[1] >= 0 else -1, r[2] + off if r[2] >= 0 else -1, r[3]]
    pop = np.stack([t0, t1])[None]
    rng = np.random.default_rng(3)
    B, h = 4, F32(0.125)
    x0 = F32(rng.normal(size=(B, 2)))
    ts = np.arange(0, 6, 0.5, dtype=F32)
    ys = F32(rng.normal(size=(B, ts.size, 2)))
    data = RolloutData(kind=abi.EVAL_SR, env_id=abi.ENV_NONE, x0=x0, ts=ts, dt0=float(h), max_fitness=1e5, ys_true=ys)
    out = oracle.rollout_population(pop, lib.table(), data, trajectories=True)
    for b in range(B):
        x = x0[b].copy()
        pred = [x.copy()]
        for s in range(int(ts[-1] / h)):
            dx = F32([x[1], F32(F32(F32(F32(1.0) - F32(x[0] * x[0])) * x[1]) - x[0])])
            x = F32(x + F32(h * dx))
            if (s + 1) % 4 == 0:
                pred.append(x.copy())
        pred = np.array(pred)
        assert (out["xs"][0, b].view(np.uint32) == pred.view(np.uint32)).all()
        mse = np.mean(np.sum((pred.astype(np.float64) - ys[b]) ** 2, axis=-1))
        np.testing.assert_allclose(out["traj_fitness"][0, b], mse, rtol=1e-5)


def test_event_terminates_and_fills_inf(oracle):
    """Latent state da/dt = a*a + 1 blows up in finite time -> cond_fn_nan fires -> the rest of the
    saved trajectory stays inf; Acrobot cost is then K + cost[0] (never successful), still finite."""
    env = Acrobot(0.0, 0.0)
    lib = util.control_library(env, abi.EVAL_DYNAMIC, state_size=1)
    ids = lib.string_to_node
    N = 8
    blow = np.tile(F32([0, -1, -1, 0]), (N, 1))
    blow[-1] = [ids["+"], N - 2, N - 5, 0]; blow[-2] = [ids["*"], N - 3, N - 4, 0]
    blow[-3] = [ids["a1"], -1, -1, 0]; blow[-4] = [ids["a1"], -1, -1, 0]; blow[-5] = [1, -1, -1, 1.0]
    read = np.tile(F32([0, -1, -1, 0]), (N, 1)); read[-1] = [1, -1, -1, 0.3]
    pop = np.stack([blow, read])[None]
    rng = np.random.default_rng(4)
    B = 3
    x0, targets = env.sample_init_states(B, rng)
    ts = np.arange(0, 10, 0.2, dtype=F32)
    data = RolloutData(kind=abi.EVAL_DYNAMIC, env_id=env.env_id, x0=x0, ts=ts, dt0=0.05, max_fitness=1e4,
                       targets=targets, params=np.ones((B, 4), F32), n_obs=4, state_size=1)
    out = oracle.rollout_population(pop, lib.table(), data, trajectories=True)
    act = out["activities"][0, :, :, 0]
    first_inf = np.argmax(~np.isfinite(act), axis=1)
    assert (first_inf > 5).all() and (first_inf < ts.size - 1).all()     # tan(t) blows up near t = pi/2 ... Euler later
    for b in range(B):
        # the save made during the terminating step is kept: the env part is still finite there, the latent
        # part already interpolates towards inf; every later save point keeps the inf pre-fill
        assert np.isinf(out["xs"][0, b, first_inf[b] + 1:]).all() and np.isfinite(out["xs"][0, b, :first_inf[b] + 1]).all()
        assert np.isinf(act[b, first_inf[b]:]).all()
    K = ts.size
    np.testing.assert_allclose(out["traj_fitness"][0], K + 0.01 * 0.3 ** 2, rtol=1e-6)
    assert out["steps"] < B * 200 * 0.9                                    # the solve really stopped early
    np.testing.assert_allclose(out["fitness"][0], K + 0.01 * 0.3 ** 2, rtol=1e-6)


def test_penalty_mean_clip(oracle):
    lib, pop, data = util.sr_case(np.random.default_rng(5), P=200, B=8)
    out = oracle.rollout_population(pop, lib.table(), data)
    tf = out["traj_fitness"]
    want = np.clip(np.where(np.isfinite(tf), tf, F32(1e5)).astype(np.float64).mean(axis=1), 0, 1e5)
    np.testing.assert_allclose(out["fitness"], want, rtol=1e-6)
    assert (~np.isfinite(tf)).any() and np.isfinite(tf).any()
    np.testing.assert_array_equal(oracle.live_count(pop), (pop[..., 0] != 0).sum((1, 2)))


# ---- Dopri5 + PIDController (what the notebooks pass) ---------------------------------------------------
def test_dopri5_accuracy_and_adaptivity(oracle):
    """Constant policy on the harmonic oscillator: the exact solution is known (scipy, 1e-12).  Fixed-step
    Dopri5 must be float32-exact at h = 0.125, the PID controller must take fewer steps at looser
    tolerances and land within a few times the tolerance, including at the dense-output save points."""
    from scipy.integrate import solve_ivp
    env = HarmonicOscillator(0.0, 0.0)
    lib = util.control_library(env, abi.EVAL_STATIC)
    rng = np.random.default_rng(2)
    B = 4
    x0, targets = env.sample_init_states(B, rng)
    omega, zeta = env.sample_params(B, "Different", None, rng)
    ts = np.arange(0, 8, 0.37, dtype=F32)          # save points that never coincide with steps
    pop = _const_policy_pop(lib, 12, 0.75)

    def run(**kw):
        data = RolloutData(kind=abi.EVAL_STATIC, env_id=env.env_id, x0=x0, ts=ts, dt0=kw.pop("dt0"), max_fitness=1e4,
                           targets=targets, params=np.stack([omega, zeta], 1), n_obs=2, **kw)
        return oracle.rollout_population(pop, lib.table(), data, trajectories=True)

    ref = []
    for b in range(B):
        A = np.array([[0, 1], [-omega[b], -zeta[b]]], dtype=np.float64)
        sol = solve_ivp(lambda t, x: A @ x + np.array([0, 0.75]), (0, float(ts[-1])), x0[b].astype(np.float64),
                        t_eval=ts.astype(np.float64), rtol=1e-12, atol=1e-12)
        ref.append(sol.y.T)
    ref = np.array(ref)
    fixed = run(dt0=0.125, solver=abi.SOLVER_DOPRI5)
    assert np.abs(fixed["xs"][0] - ref).max() < 3e-5
    loose = run(dt0=0.05, solver=abi.SOLVER_DOPRI5, controller=abi.CONTROLLER_PID, rtol=1e-3, atol=1e-3, dtmin=0.001)
    tight = run(dt0=0.05, solver=abi.SOLVER_DOPRI5, controller=abi.CONTROLLER_PID, rtol=1e-6, atol=1e-6, dtmin=0.001)
    assert loose["steps"] < tight["steps"] < fixed["steps"]
    assert np.abs(loose["xs"][0] - ref).max() < 2e-2 and np.abs(tight["xs"][0] - ref).max() < 5e-5
    # max_steps counts rejected steps too and leaves the tail inf (throw=False)
    short = run(dt0=0.05, solver=abi.SOLVER_DOPRI5, controller=abi.CONTROLLER_PID, rtol=1e-6, atol=1e-6, dtmin=0.001,
                max_steps=8)
    assert np.isinf(short["xs"][0, :, -1]).all() and np.isfinite(short["xs"][0, :, 0]).all()
    assert (short["fitness"] == 1e4).all()


def test_dynamic_acrobot_against_numpy_restatement(oracle):
    """The C oracle's DynamicPolicy/Acrobot path (obs with noise and wrap, readout on [0,a,0], drift, state
    equation on [y,a,u], Euler, SaveAt interpolation, post-pass, Acrobot cost) against an independent NumPy
    float32 restatement written from the reference sources.  Short horizon, libm-vs-NumPy tolerance."""
    rng = np.random.default_rng(17)
    env = Acrobot(0.0, 0.1)
    lib = util.control_library(env, abi.EVAL_DYNAMIC, state_size=2)
    fns = np_reference.node_functions(lib)
    P, B = 12, 3
    pop = sample_population(rng, P, lib, 4, 30)
    data = util.control_data(env, abi.EVAL_DYNAMIC, rng, B, 2.0, 0.2, 0.07, 2)   # dt0 does not divide the save spacing
    oracle.set_math(oracle.MATH_LIBM)
    try:
        out = oracle.rollout_population(pop, lib.table(), data, trajectories=True)
    finally:
        oracle.set_math(oracle.MATH_CONTRACT)
    checked = 0
    for p in range(P):
        for b in range(B):
            key = data.obs_keys[b]

            def noise_at(t, key=key):
                fk = oracle.fold_in(key, int(np.float32(t).view(np.uint32)))
                return (oracle.normal(fk, 4, True) * F32(0.1)).astype(F32)
            xs, acts, us, fit = np_reference.dynamic_acrobot_rollout(pop[p], fns, 2, data.x0[b], data.ts, data.dt0,
                                                                    data.params[b], noise_at)
            ref_xs, ref_us = out["xs"][p, b], out["us"][p, b, :, 0]
            fin = np.isfinite(xs).all(axis=1)
            np.testing.assert_array_equal(fin, np.isfinite(ref_xs).all(axis=1))
            np.testing.assert_allclose(ref_xs[fin], xs[fin], rtol=2e-4, atol=2e-4)
            np.testing.assert_allclose(out["activities"][p, b][fin], acts[fin], rtol=2e-3, atol=2e-3)
            ok = np.isfinite(us) & np.isfinite(ref_us)
            np.testing.assert_allclose(ref_us[ok], us[ok], rtol=2e-3, atol=2e-3)
            if np.isfinite(fit) and np.isfinite(out["traj_fitness"][p, b]):
                np.testing.assert_allclose(out["traj_fitness"][p, b], fit, rtol=1e-3)
            checked += 1
    assert checked == P * B


@pytest.mark.filterwarnings("ignore::RuntimeWarning")
def test_reactor_against_numpy_restatement(oracle):
    """StirredTankReactor drift (reactor.py:78-87, Arrhenius term, clipped control) and cost (:92-96) with a
    constant policy: independent NumPy float32 Euler vs the oracle (libm mode)."""
    from multitreegp_b200.environments.control_environments import StirredTankReactor
    env = StirredTankReactor(0.0, 0.0)
    lib = util.control_library(env, abi.EVAL_STATIC)
    rng = np.random.default_rng(23)
    B, h, u0 = 4, F32(0.002), F32(120.0)
    x0, targets = env.sample_init_states(B, rng)
    params = np.stack(env.sample_params(B, "Different", None, rng), 1)
    ts = np.arange(0, 0.4, 0.02, dtype=F32)
    data = RolloutData(kind=abi.EVAL_STATIC, env_id=env.env_id, x0=x0, ts=ts, dt0=float(h), max_fitness=1e4,
                       targets=targets, params=params, n_obs=3)
    oracle.set_math(oracle.MATH_LIBM)
    try:
        out = oracle.rollout_population(_const_policy_pop(lib, 10, float(u0)), lib.table(), data, trajectories=True)
    finally:
        oracle.set_math(oracle.MATH_CONTRACT)
    for b in range(B):
        Vol, Cp, dHr, UA, q, Tf, Tcf, Volc = [F32(v) for v in params[b]]
        k = lambda T: F32(7.2e10) * np.exp(F32(-72750 / 8.314) / T)
        x = x0[b].astype(F32)
        t, saved, kk = F32(0), [], 0
        tprev, tnext, t1 = F32(ts[0]), F32(ts[0] + h), F32(ts[-1])
        while tprev < t1:
            Tc, T, c = x
            control = np.clip(u0, F32(0), F32(300))
            dc = (q / Vol) * (F32(1.0) - c) - k(T) * c
            dT = (q / Vol) * (Tf - T) + (-dHr / Cp) * k(T) * c + (UA / Vol / Cp) * (Tc - T)
            dTc = (control / Volc) * (Tcf - Tc) + (UA / Volc / Cp) * (T - Tc)
            x1 = (x + (tnext - tprev) * np.array([dTc, dT, dc], dtype=F32)).astype(F32)
            while kk < len(ts) and ts[kk] <= tnext:
                w = F32((ts[kk] - tprev) / (tnext - tprev))
                saved.append((x + w * (x1 - x)).astype(F32)); kk += 1
            x, tprev = x1, min(tnext, t1)
            nn = F32(tnext + h)
            tnext = t1 if nn > F32(t1 - F32(1e-6)) else nn
            if not np.isfinite(x1).all():          # cond_fn_nan (reactor.py:98-99): the solve stops here
                break
        saved += [np.full(3, np.inf, F32)] * (len(ts) - len(saved))
        xs = np.array(saved)
        fin = np.isfinite(xs).all(axis=1)
        np.testing.assert_array_equal(fin, np.isfinite(out["xs"][0, b]).all(axis=1))
        np.testing.assert_allclose(out["xs"][0, b][fin], xs[fin], rtol=1e-5, atol=1e-4)
        if fin.all():
            cost = np.sum(0.01 * (xs[:, 1].astype(np.float64) - targets[b, 0]) ** 2 + 1e-4 * float(u0) ** 2)
            np.testing.assert_allclose(out["traj_fitness"][0, b], cost, rtol=1e-4)
        else:
            assert not np.isfinite(out["traj_fitness"][0, b])
