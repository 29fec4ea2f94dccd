"""Host-side mirror of the reference interface, without a GPU: evaluator constructors with the reference's
signatures (SR_evaluator.py:21, feedforward_evaluate.py:11, dynamic_evaluate.py:11), the data-tuple layouts
(feedforward_evaluate.py:62), solver / step-size-controller recognition, and the error paths."""
import numpy as np
import pytest

from multitreegp_b200 import abi
from multitreegp_b200.environments.control_environments import Acrobot, HarmonicOscillator, StirredTankReactor
from multitreegp_b200.evaluators import SR_evaluator, dynamic_evaluate, feedforward_evaluate
from multitreegp_b200.solvers import ConstantStepSize, Dopri5, Euler, PIDController
from tests import util


def _control_tuple(env, rng, B, K):
    x0, targets = env.sample_init_states(B, rng)
    ts = np.arange(K, dtype=np.float32) * np.float32(0.2)
    keys = util.random_keys(rng, B)
    return (x0, ts, targets, keys, keys, env.sample_params(B, "Different", ts, rng))


@pytest.mark.parametrize("env_cls,n_params", [(Acrobot, 4), (HarmonicOscillator, 2), (StirredTankReactor, 8)])
def test_control_data_tuple_becomes_rollout_data(env_cls, n_params):
    rng = np.random.default_rng(1)
    env = env_cls(0.05, 0.1)
    B, K = 6, 11
    data = _control_tuple(env, rng, B, K)
    ev = feedforward_evaluate.Evaluator(env, 0.05)
    rd = ev.rollout_data(data)
    assert rd.kind == abi.EVAL_STATIC and rd.env_id == env.env_id
    assert rd.x0.shape == (B, env.n_var) and rd.ts.shape == (K,) and rd.params.shape == (B, n_params)
    assert rd.obs_keys.dtype == np.uint32 and rd.obs_keys.shape == (B, 2)
    assert rd.n_obs == env.n_obs and rd.state_size == 0 and rd.max_steps == 16 ** 4 and rd.dt0 == 0.05
    assert rd.obs_noise == pytest.approx(0.1) and rd.max_fitness == 1e4
    assert rd.solver == abi.SOLVER_EULER and rd.controller == abi.CONTROLLER_CONSTANT  # the class defaults
    ev2 = dynamic_evaluate.Evaluator(env, 3, 0.025, max_steps=500)
    rd2 = ev2.rollout_data(data)
    assert rd2.kind == abi.EVAL_DYNAMIC and rd2.state_size == 3 and rd2.max_steps == 500 and rd2.dt0 == 0.025
    with pytest.raises(ValueError):  # a params pytree of the wrong length
        ev.rollout_data(data[:5] + (data[5][:-1],))


def test_solver_arguments_as_in_the_notebooks():
    env = HarmonicOscillator(0.0, 0.0)
    ev = feedforward_evaluate.Evaluator(env, 0.05, solver=Dopri5(),
                                        stepsize_controller=PIDController(atol=1e-4, rtol=1e-4, dtmin=0.001), max_steps=1000)
    rd = ev.rollout_data(_control_tuple(env, np.random.default_rng(2), 4, 5))
    assert (rd.solver, rd.controller) == (abi.SOLVER_DOPRI5, abi.CONTROLLER_PID)
    assert rd.rtol == pytest.approx(1e-4) and rd.atol == pytest.approx(1e-4) and rd.dtmin == pytest.approx(1e-3)
    assert rd.max_steps == 1000
    # explicit defaults are the defaults
    ev = SR_evaluator.Evaluator(solver=Euler(), dt0=0.01, stepsize_controller=ConstantStepSize())
    assert ev._solver_settings.get("solver", abi.SOLVER_EULER) == abi.SOLVER_EULER
    # an adaptive controller needs an error estimate; unknown solvers are refused at construction
    with pytest.raises((ValueError, NotImplementedError)):
        feedforward_evaluate.Evaluator(env, 0.05, solver=Euler(), stepsize_controller=PIDController(atol=1e-4, rtol=1e-4))
    with pytest.raises(NotImplementedError):
        feedforward_evaluate.Evaluator(env, 0.05, solver=type("Tsit5", (), {})())


def test_sr_data_tuple_and_unbound_evaluator():
    from bench_support.synthetic import van_der_pol_data
    rng = np.random.default_rng(3)
    x0, ts, ys, keys = van_der_pol_data(rng, 5, 2.0, 0.2)
    ev = SR_evaluator.Evaluator()
    rd = ev.rollout_data((x0, ts, ys, keys))
    assert rd.kind == abi.EVAL_SR and rd.env_id == abi.ENV_NONE and rd.max_fitness == 1e5
    assert rd.ys_true.shape == (5, ts.size, 2) and rd.x0.shape == (5, 2)
    with pytest.raises(RuntimeError):  # node ids are assigned by GeneticProgramming.__init__: bind() first
        ev._table()
    ev.bind([("+", 2, 0.5), ("*", 2, 0.5)], [["x0", "x1"]], [2])
    t = ev._table()
    assert t.n_ids == 2 + 2 + 2


def test_environment_classes_mirror_the_reference_constructors():
    """acrobot.py:7-27, harmonic_oscillator.py:8-56, reactor.py:7-65: sizes, initial-state bounds, parameter pytrees."""
    rng = np.random.default_rng(4)
    ts = np.arange(0, 2, 0.2, dtype=np.float32)
    acro = Acrobot(0.0, 0.1)
    assert (acro.n_var, acro.n_control, acro.n_targets, acro.n_obs) == (4, 1, 0, 4)
    x0, targets = acro.sample_init_states(64, rng)
    assert x0.shape == (64, 4) and x0.dtype == np.float32 and np.abs(x0).max() <= 0.1   # U(-0.1, 0.1) (acrobot.py:18-22)
    assert targets.shape == (64, 0)                                                     # zero-size target is legal
    p = acro.sample_params(8, "Constant", ts, rng)
    assert len(p) == 4 and all(np.all(a == 1.0) and a.shape == (8,) for a in p)         # l1 = l2 = m1 = m2 = 1
    assert Acrobot(0.0, 0.0, n_obs=2).n_obs == 2                                        # partial observation

    ho = HarmonicOscillator(0.0, 0.0)
    assert (ho.n_var, ho.n_control, ho.n_targets) == (2, 1, 1)
    om, ze = ho.sample_params(16, "Constant", ts, rng)
    assert np.all(om == 1.0) and np.all(ze == 0.0)
    om, ze = ho.sample_params(256, "Different", ts, rng)
    assert 0.0 <= om.min() and om.max() <= 2.0 and 0.0 <= ze.min() and ze.max() <= 1.5
    with pytest.raises(NotImplementedError):
        ho.sample_params(4, "Switch", ts, rng)      # time-varying parameters: out of scope, refused loudly
    x0, targets = ho.sample_init_states(32, rng)
    assert x0.shape == (32, 2) and targets.shape == (32, 1) and np.abs(targets).max() <= 3.0

    cstr = StirredTankReactor(0.0, 0.0)
    assert (cstr.n_var, cstr.n_control, cstr.n_targets) == (3, 1, 1)
    p = cstr.sample_params(8, "Constant", ts, rng)
    assert len(p) == 8 and all(a.shape == (8,) for a in p)
    for env in (acro, ho, cstr):
        assert env.n_params == len(env.sample_params(2, "Constant", ts, rng))
