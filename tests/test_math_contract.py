"""Accuracy of the numerical contract (include/mtgp_math.h), compiled into the oracle, against
float64 references.  The GPU side compiles the very same header; bit-equality between the two is
asserted by the -m gpu parity tests."""
import numpy as np
import pytest


def _ulp_err(got, ref64):
    got = got.astype(np.float64)
    ulp = np.spacing(np.abs(ref64).astype(np.float32)).astype(np.float64)
    ok = np.isfinite(ref64) & (np.abs(ref64) > 1e-37) & (np.abs(ref64) < 3e38)
    return float((np.abs(got - ref64)[ok] / ulp[ok]).max())


def test_sin_cos_fast_path(oracle):
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-10, 10, 400000), rng.uniform(-1e5, 1e5, 400000),
                        rng.normal(size=100000) * 1e-3, np.arange(-400, 400) * (np.pi / 4)]).astype(np.float32)
    assert _ulp_err(oracle.math_apply("sin", x), np.sin(x.astype(np.float64))) < 2.0
    assert _ulp_err(oracle.math_apply("cos", x), np.cos(x.astype(np.float64))) < 2.0


def test_sin_cos_payne_hanek(oracle):
    import mpmath as mp
    mp.mp.prec = 300
    rng = np.random.default_rng(1)
    x = (rng.uniform(1, 2, 4000) * 2.0 ** rng.integers(17, 127, 4000)).astype(np.float32)
    x *= rng.choice([-1, 1], x.size).astype(np.float32)
    x = np.concatenate([x, np.float32([1e5, 100000.01, 3.4e38, -3.4e38])])
    s = np.array([float(mp.sin(mp.mpf(float(v)))) for v in x])
    c = np.array([float(mp.cos(mp.mpf(float(v)))) for v in x])
    assert _ulp_err(oracle.math_apply("sin", x), s) < 2.0
    assert _ulp_err(oracle.math_apply("cos", x), c) < 2.0


def test_exp_log_tanh_log1p(oracle):
    rng = np.random.default_rng(2)
    x = rng.uniform(-103, 88.7, 500000).astype(np.float32)
    assert _ulp_err(oracle.math_apply("exp", x), np.exp(x.astype(np.float64))) < 1.01
    x = np.concatenate([np.exp(rng.uniform(-87, 88, 400000)), rng.uniform(0.5, 2, 200000),
                        rng.uniform(1e-45, 1e-38, 1000)]).astype(np.float32)
    assert _ulp_err(oracle.math_apply("log", x), np.log(x.astype(np.float64))) < 1.01
    x = rng.uniform(-12, 12, 500000).astype(np.float32)
    assert _ulp_err(oracle.math_apply("tanh", x), np.tanh(x.astype(np.float64))) < 2.5
    x = np.concatenate([-rng.uniform(0, 1, 300000) ** 2, rng.uniform(-1e-3, 1e-3, 50000)]).astype(np.float32)
    assert _ulp_err(oracle.math_apply("log1p", x), np.log1p(x.astype(np.float64))) < 3.0


def test_special_values(oracle):
    inf, nan = np.float32(np.inf), np.float32(np.nan)
    f = lambda name, v: oracle.math_apply(name, np.float32([v]))[0]
    assert np.isnan(f("sin", inf)) and np.isnan(f("cos", -inf)) and np.isnan(f("sin", nan))
    assert f("sin", 0.0) == 0 and f("cos", 0.0) == 1
    assert f("exp", inf) == inf and f("exp", -inf) == 0 and np.isnan(f("exp", nan)) and f("exp", 0.0) == 1
    assert f("exp", 89.0) == inf and f("exp", -110.0) == 0
    assert f("log", 0.0) == -inf and np.isnan(f("log", -1.0)) and f("log", inf) == inf and f("log", 1.0) == 0
    assert f("tanh", 50.0) == 1 and f("tanh", -50.0) == -1 and np.isnan(f("tanh", nan))
    assert f("log1p", -1.0) == -inf
    assert f("erfinv", 1.0) == inf and f("erfinv", -1.0) == -inf and f("erfinv", 0.0) == 0


def test_erfinv_matches_its_definition(oracle):
    """Giles' polynomial in w = -log1p(-x*x) (XLA's ErfInv32).  Near |x| = 1 the float32 x*x makes it
    deviate from the true erfinv by design, so compare the central region with scipy and the whole
    range with a float64 evaluation of the same formula."""
    from scipy.special import erfinv
    rng = np.random.default_rng(3)
    x = rng.uniform(-0.9, 0.9, 200000).astype(np.float32)
    got = oracle.math_apply("erfinv", x)
    np.testing.assert_allclose(got, erfinv(x.astype(np.float64)), rtol=2e-6, atol=1e-7)
    x = rng.uniform(-1, 1, 200000).astype(np.float32)
    w = -np.log1p(-(x * x).astype(np.float64))
    c1 = [2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06, 0.00021858087, -0.00125372503,
          -0.00417768164, 0.246640727, 1.50140941]
    c2 = [-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844, 0.00573950773, -0.0076224613,
          0.00943887047, 1.00167406, 2.83297682]
    p1 = np.polyval(c1, w - 2.5)
    p2 = np.polyval(c2, np.sqrt(np.maximum(w, 5)) - 3)
    ref = np.where(w < 5, p1, p2) * x
    np.testing.assert_allclose(oracle.math_apply("erfinv", x), ref, rtol=3e-5, atol=1e-7)


def test_python_remainder_is_exact(oracle):
    rng = np.random.default_rng(4)
    x = np.concatenate([rng.uniform(-50, 50, 200000), rng.normal(size=50000) * 1e6,
                        rng.normal(size=20000) * 1e30, rng.normal(size=1000) * 1e-40]).astype(np.float32)
    y = np.full_like(x, np.float32(2 * np.pi))
    got = oracle.math_apply("remainder", x, y)
    ref = np.mod(x.astype(np.float64), y.astype(np.float64))  # exact in float64 for these magnitudes? use fmod
    ref = np.fmod(x.astype(np.float64), y.astype(np.float64))
    ref = np.where((ref != 0) & (ref < 0), (ref.astype(np.float32) + y).astype(np.float64), ref)
    small = np.abs(x) < 1e15  # float64 fmod is exact; the float32 result of fmod is exactly representable
    np.testing.assert_array_equal(got[small], ref[small].astype(np.float32))
    # general y, including negative divisors and tiny / huge ratios
    y = (rng.normal(size=x.size) * 10 ** rng.uniform(-3, 3, x.size)).astype(np.float32)
    got = oracle.math_apply("remainder", x, y)
    r = np.fmod(x.astype(np.float64), y.astype(np.float64)).astype(np.float32)
    ref = np.where((r != 0) & ((r < 0) != (y < 0)), r + y, r).astype(np.float32)
    ok = np.abs(x.astype(np.float64) / y.astype(np.float64)) < 2 ** 52
    np.testing.assert_array_equal(got[ok], ref[ok])


def test_libm_mode_is_close_to_contract(oracle):
    rng = np.random.default_rng(5)
    x = rng.uniform(-20, 20, 100000).astype(np.float32)
    a = {n: oracle.math_apply(n, x) for n in ("sin", "cos", "tanh")}
    oracle.set_math(oracle.MATH_LIBM)
    try:
        for n, v in a.items():
            np.testing.assert_allclose(oracle.math_apply(n, x), v, rtol=3e-7, atol=2e-7)
    finally:
        oracle.set_math(oracle.MATH_CONTRACT)
