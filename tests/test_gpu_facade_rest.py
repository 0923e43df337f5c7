"""GPU parity for the rest of the reference's C facade: elementwise math, reorganisation,
reductions.  Checker: oracle/dsp_oracle.py (NumPy restatement; complex functions follow num-complex 0.4).

Tolerances (written next to each assert): real libm functions <= 4 ulp of T (pow-like <= 8 ulp); complex
functions |err| <= 16 eps_T * max(1, |ref|) (their formulas cancel, so ulp of the result is not meaningful);
data movement and fmod are bit-exact; sums within 1e-6 (f32 results) / 1e-13 (f64, prec) of sum(|x|)."""
import ctypes
import math

import numpy as np
import pytest

import basic_dsp_b200 as bd
from basic_dsp_b200 import DspVec
from oracle import dsp_oracle as o

pytestmark = pytest.mark.gpu

DT = [np.float32, np.float64]


def rand_c(rng, n, dtype, lo=-2.0, hi=2.0):
    ct = np.complex64 if dtype == np.float32 else np.complex128
    return (rng.uniform(lo, hi, n) + 1j * rng.uniform(lo, hi, n)).astype(ct)


REAL_FUNCS = [("sin", -10, 10), ("cos", -10, 10), ("tan", -1.5, 1.5), ("asin", -1, 1), ("acos", -1, 1), ("atan", -10, 10),
              ("sinh", -5, 5), ("cosh", -5, 5), ("tanh", -5, 5), ("asinh", -10, 10), ("acosh", 1, 10), ("atanh", -0.99, 0.99),
              ("sqrt", 0, 100), ("square", -10, 10), ("ln", 0.01, 100), ("exp", -10, 10), ("abs", -10, 10)]


@pytest.mark.parametrize("dtype", DT)
def test_real_math(dtype):  # trigonometry_and_powers.rs:198-377, tests/real_test.rs pattern (op vs scalar closure)
    rng = np.random.default_rng(1)
    n = 20001
    for name, lo, hi in REAL_FUNCS:
        x = rng.uniform(lo, hi, n).astype(dtype)
        got = DspVec(x).math(name).to_numpy()
        ref = o.real_math(name, x, dtype)
        assert o.ulp_diff(got, ref, dtype).max() <= 4, name
    x = rng.uniform(0.1, 10, n).astype(dtype)
    for name, arg in [("powf", 2.5), ("root", 3.0), ("log", 10.0), ("expf", 3.0)]:
        got = DspVec(x).math(name, arg).to_numpy()
        assert o.ulp_diff(got, o.real_math(name, x, dtype, arg), dtype).max() <= 8, name
    for name, arg, base in [("powf_approx", 2.5, "powf"), ("log_approx", 10.0, "log"), ("expf_approx", 3.0, "expf")]:
        got = DspVec(x).math(name, arg).to_numpy()
        assert np.max(np.abs(got / o.real_math(base, x, dtype, arg) - 1)) < 1e-2      # the reference promises ~1 %
    for name, base in [("ln_approx", "ln"), ("exp_approx", "exp"), ("sin_approx", "sin"), ("cos_approx", "cos")]:
        got = DspVec(x).math(name).to_numpy()
        assert np.max(np.abs(got - o.real_math(base, x, dtype))) < 1e-2 * np.max(np.abs(o.real_math(base, x, dtype)))
    # wrap = fmod: bit-exact
    x = rng.uniform(-50, 50, n).astype(dtype)
    assert np.array_equal(DspVec(x).math("wrap", 2 * math.pi).to_numpy(), o.real_math("wrap", x, dtype, 2 * math.pi))
    # complex vectors are rejected by the real-only operations
    assert DspVec(rand_c(rng, 8, dtype)).result_code_of("abs") == -1
    assert DspVec(rand_c(rng, 8, dtype)).result_code_of("wrap", 1.0) == -1


@pytest.mark.parametrize("dtype", DT)
def test_complex_math(dtype):  # num-complex 0.4 formulas; complex_test.rs pattern
    rng = np.random.default_rng(2)
    n = 20001
    eps = np.finfo(dtype).eps
    z = rand_c(rng, n, dtype)
    for name in ("sin", "cos", "tan", "sinh", "cosh", "tanh", "asin", "acos", "atan", "asinh", "acosh", "atanh", "sqrt",
                 "square", "ln", "exp"):
        zz = z
        if name == "tan":      # poles at re = +-pi/2, im = 0: the formula's denominator cancels there (in the reference too)
            zz = (np.clip(z.real, -1.2, 1.2) + 1j * z.imag).astype(z.dtype)
        if name == "tanh":     # poles at im = +-pi/2, re = 0
            zz = (z.real + 1j * np.clip(z.imag, -1.2, 1.2)).astype(z.dtype)
        got = DspVec(zz).math(name).to_numpy()
        ref = o.complex_math(name, zz, dtype)
        err = np.abs(got.astype(np.complex128) - ref.astype(np.complex128)) / np.maximum(1.0, np.abs(ref))
        assert err.max() <= 16 * eps, (name, err.max())
    for name, arg in [("powf", 2.5), ("root", 3.0), ("log", 10.0), ("expf", 3.0)]:
        got = DspVec(z).math(name, arg).to_numpy()
        ref = o.complex_math("powf" if name == "root" else name, z, dtype, dtype(1) / dtype(arg) if name == "root" else arg)
        err = np.abs(got.astype(np.complex128) - ref.astype(np.complex128)) / np.maximum(1.0, np.abs(ref))
        assert err.max() <= 32 * eps, (name, err.max())
    # exact branches of Complex::sqrt
    s = np.array([complex(4, 0), complex(-4, 0.0), complex(-4, -0.0), complex(0, 2), complex(0, -2)], dtype=z.dtype)
    got = DspVec(s).math("sqrt").to_numpy()
    assert np.allclose(got, [2, 2j, -2j, 1 + 1j, 1 - 1j], atol=4 * eps)


@pytest.mark.parametrize("dtype", DT)
def test_unwrap_diff_cumsum(dtype):  # real_ops.rs:266-288, diff_sum.rs:63-122
    rng = np.random.default_rng(3)
    for n in (1, 2, 1000, 1024, 1025, 5000):
        ph = np.cumsum(rng.uniform(-2, 2, n))
        x = np.fmod(ph, 2 * math.pi).astype(dtype)
        got = DspVec(x).math("unwrap", 2 * math.pi).to_numpy()
        assert np.array_equal(got, o.unwrap(x, 2 * math.pi, dtype)), n          # bit-exact (sequential semantics)
    for cplx in (False, True):
        for n in (2, 3, 4097, 100000):
            x = rand_c(rng, n, dtype) if cplx else rng.uniform(-10, 10, n).astype(dtype)
            assert np.array_equal(DspVec(x).math("diff").to_numpy(), o.diff(x))
            assert np.array_equal(DspVec(x).math("diff_with_start").to_numpy(), o.diff(x, with_start=True))
            got = DspVec(x).math("cum_sum").to_numpy()
            ref = np.cumsum(x.astype(np.complex128 if cplx else np.float64))
            # running sums: parallel scan vs sequential order -> tolerance relative to the running sum of |x|
            bound = np.finfo(dtype).eps * 8 * np.cumsum(np.abs(x).astype(np.float64)) + 1e-30
            assert np.all(np.abs(got - ref) <= bound * max(1.0, math.log2(n)))
    # doc examples (diff_sum.rs:13-60)
    assert DspVec(np.array([2.0, 3.0, 2.0, 6.0], dtype=dtype)).math("diff").to_numpy().tolist() == [1.0, -1.0, 4.0]
    assert DspVec(np.array([2.0, 3.0, 2.0, 6.0], dtype=dtype)).math("diff_with_start").to_numpy().tolist() == [2.0, 1.0, -1.0, 4.0]
    assert DspVec(np.array([2.0, 1.0, -1.0, 4.0], dtype=dtype)).math("cum_sum").to_numpy().tolist() == [2.0, 3.0, 2.0, 6.0]


@pytest.mark.parametrize("dtype", DT)
def test_smaller_vector_ops(dtype):  # elementary.rs:457-517,601-639
    rng = np.random.default_rng(4)
    for cplx in (False, True):
        x = rand_c(rng, 6000, dtype) if cplx else rng.uniform(-10, 10, 6000).astype(dtype)
        w = rand_c(rng, 12, dtype) if cplx else rng.uniform(1, 10, 12).astype(dtype)
        for op in ("add", "sub", "mul", "div"):
            got = getattr(DspVec(x), op + "_smaller")(DspVec(w)).to_numpy()
            ref = o.binary_smaller(op, x, w, dtype)
            assert o.ulp_diff(got.view(dtype), np.asarray(ref).view(dtype), dtype).max() <= (4 if op == "div" else 0), op
        assert DspVec(x).result_code_of("mul_smaller_vector", DspVec(w[:7])) == 7     # InvalidArgumentLength


@pytest.mark.parametrize("dtype", DT)
def test_real_imag_mag_phase_split_merge(dtype):  # complex_to_real.rs:674-770, data_reorganization.rs:484-557
    rng = np.random.default_rng(5)
    z = rand_c(rng, 3000, dtype)
    v = DspVec(z)
    re, im = DspVec(np.zeros(1, dtype=dtype)), DspVec(np.zeros(1, dtype=dtype))
    assert v.get_real_imag(re, im) == 9                                             # convert_void(Ok) (Q8)
    assert np.array_equal(re.to_numpy(), z.real) and np.array_equal(im.to_numpy(), z.imag)
    back = DspVec(np.zeros(2, dtype=z.dtype)).set_real_imag(re, im)
    assert np.array_equal(back.to_numpy(), z)
    mag, ph = np.abs(z).astype(dtype), np.angle(z).astype(dtype)
    got = DspVec(np.zeros(2, dtype=z.dtype)).set_mag_phase(DspVec(mag), DspVec(ph)).to_numpy()
    ref = o.set_mag_phase(mag, ph, dtype)
    assert o.ulp_diff(got.view(dtype), ref.view(dtype), dtype).max() <= 4
    assert DspVec(np.zeros(2, dtype=z.dtype)).result_code_of("set_real_imag", DspVec(mag), DspVec(ph[:5])) == 7
    for cplx in (False, True):
        x = rand_c(rng, 3000, dtype) if cplx else rng.uniform(-10, 10, 3000).astype(dtype)
        for parts in (1, 3, 5):
            targets = [DspVec(np.zeros(2, dtype=x.dtype)) for _ in range(parts)]
            assert DspVec(x).split_into(targets) == 9
            ref = o.split_into(x, parts)
            for t, r in zip(targets, ref):
                assert np.array_equal(t.to_numpy(), r)
            merged = DspVec(np.zeros(2, dtype=x.dtype)).merge(targets)
            assert np.array_equal(merged.to_numpy(), x)
        assert DspVec(x).split_into([DspVec(np.zeros(2, dtype=x.dtype)) for _ in range(7)]) == 7   # 3000 % 7 != 0


@pytest.mark.parametrize("dtype", DT)
def test_interpolate_hermite(dtype):  # real_interpolation.rs:196-228 + oracle parity
    got = DspVec(np.array([-1.0, -2.0, -1.0, 0.0, 1.0, 3.0, 4.0], dtype=dtype), domain=bd.FREQ).interpolate_hermite(4.0, 0.0).to_numpy()
    exp = [-1.0000, -1.4375, -1.7500, -1.9375, -2.0000, -1.8906, -1.6250, -1.2969, -1.0000, -0.7500, -0.5000, -0.2500, 0.0,
           0.2344, 0.4583, 0.7031, 1.0000, 1.4375, 2.0000, 2.5625, 3.0000, 3.3203, 3.6042, 3.8359, 4.0]
    assert len(got) == 25 and np.max(np.abs(got[4:-4] - np.array(exp)[4:-4])) < 6e-2      # hermit_spline_test
    got = DspVec(np.array([-3.0, -2.0, -1.0, 0.0, 1.0, 2.0, 3.0], dtype=dtype)).interpolate_hermite(3.0, 0.0).to_numpy()
    assert np.max(np.abs(got - np.linspace(-3, 3, 19))) < 5e-3                          # hermit_spline_test_linear_increment
    rng = np.random.default_rng(6)
    for n, F, d in [(1000, 4.0, 0.0), (777, 3.0, 0.0), (500, 2.5, 0.0)]:  # (a delay > 0 makes the reference index out of bounds)
        x = rng.uniform(-10, 10, n).astype(dtype)
        got = DspVec(x).interpolate_hermite(F, d).to_numpy()
        ref = o.interpolate_hermite(x, F, d, dtype)
        assert len(got) == len(ref)
        assert o.ulp_diff(got, ref, dtype).max() <= 0, (n, F, d)                        # every operation rounded in T: bit-exact


@pytest.mark.parametrize("dtype", DT)
def test_custom_windows(dtype):  # facade32.rs:1030-1139
    rng = np.random.default_rng(7)
    n = 1001
    hamming = lambda i, p: 0.54 - 0.46 * math.cos(2 * math.pi * i / (p - 1))
    ramp = lambda i, p: 1.0 + i / p
    x = rand_c(rng, n, dtype)
    got = DspVec(x).custom_window("apply_custom_window", hamming).to_numpy()
    ref = DspVec(x).apply_window(bd.HAMMING).to_numpy()
    assert o.rel_l2(got, ref) < 1e-6
    w = np.array([dtype(ramp(i, n)) for i in range(n)])
    got = DspVec(x).custom_window("apply_custom_window", ramp, is_symmetric=False).to_numpy()
    assert o.rel_l2(got, x * w) < (1e-6 if dtype == np.float32 else 1e-14)
    got = DspVec(x).custom_window("unapply_custom_window", ramp, is_symmetric=False).to_numpy()
    assert o.rel_l2(got, x / w) < (1e-6 if dtype == np.float32 else 1e-14)
    got = DspVec(x).custom_window("windowed_custom_fft", hamming).to_numpy()
    assert o.rel_l2(got, DspVec(x).windowed_fft(bd.HAMMING).to_numpy()) < 1e-5
    X = DspVec(x).fft().to_numpy()
    got = DspVec(X, domain=bd.FREQ).custom_window("windowed_custom_ifft", hamming).to_numpy()
    assert o.rel_l2(got, DspVec(X, domain=bd.FREQ).windowed_ifft(bd.HAMMING).to_numpy()) < 1e-5
    r = rng.uniform(-10, 10, n).astype(dtype)
    got = DspVec(r).custom_window("windowed_custom_sfft", hamming).to_numpy()
    assert o.rel_l2(got, DspVec(r).windowed_sfft(bd.HAMMING).to_numpy()) < 1e-5


@pytest.mark.parametrize("dtype", DT)
def test_map_callbacks(dtype):  # mapping.rs:46-266
    x = np.arange(100, dtype=dtype)
    got = DspVec(x).map_inplace(lambda v, i: 2 * v + i).to_numpy()
    assert np.array_equal(got, 3 * x)
    L = bd.lib()
    s = "32" if dtype == np.float32 else "64"
    keep = []

    def mapper(v, i):
        keep.append(ctypes.c_double(v * (i + 1)))
        return ctypes.addressof(keep[-1])

    def aggr(a, b):
        keep.append(ctypes.c_double(ctypes.c_double.from_address(a).value + ctypes.c_double.from_address(b).value))
        return ctypes.addressof(keep[-1])

    v = DspVec(x)
    r = getattr(L, "map_aggregate_real" + s)(v._h, getattr(L, "MapAggregateRealFn" + s)(mapper), getattr(L, "AggregateFn" + s)(aggr))
    assert r.result_code == 0
    assert ctypes.c_double.from_address(r.result).value == float(np.sum(x.astype(np.float64) * (np.arange(100) + 1)))
    c = DspVec(np.ones(4, dtype=np.complex64 if dtype == np.float32 else np.complex128))
    r = getattr(L, "map_aggregate_real" + s)(c._h, getattr(L, "MapAggregateRealFn" + s)(mapper), getattr(L, "AggregateFn" + s)(aggr))
    assert r.result_code == 4                                                           # InputMustBeReal


# --------------------------------------------------------------------------------------------------
# reductions
# --------------------------------------------------------------------------------------------------
def _close(a, b, scale, tol):
    return abs(complex(a) - complex(b)) <= tol * scale


@pytest.mark.parametrize("dtype", DT)
def test_sums_and_dot_products(dtype):  # statistics.rs:440-560, dot_products.rs:67-345
    rng = np.random.default_rng(8)
    tol = 1e-6 if dtype == np.float32 else 1e-13
    for n in (1, 255, 4096, 1000003):
        for cplx in (False, True):
            x = rand_c(rng, n, dtype, -10, 10) if cplx else rng.uniform(-10, 10, n).astype(dtype)
            w = rand_c(rng, n, dtype, -10, 10) if cplx else rng.uniform(-10, 10, n).astype(dtype)
            wide = np.complex128 if cplx else np.float64
            xs = x.astype(wide)
            sabs = float(np.sum(np.abs(xs)))
            sabs2 = float(np.sum(np.abs(xs) ** 2))
            v = DspVec(x)
            exact_sum = complex(math.fsum(xs.real), math.fsum(xs.imag)) if cplx else math.fsum(xs)
            sq = xs * xs
            exact_sq = complex(math.fsum(sq.real), math.fsum(sq.imag)) if cplx else math.fsum(sq)
            assert _close(v.sum(), exact_sum, sabs, tol)
            assert _close(v.sum(squared=True), exact_sq, sabs2, tol)
            assert _close(v.sum(prec=True), exact_sum, sabs, 1e-14)
            assert _close(v.sum(prec=True, squared=True), exact_sq, sabs2, 1e-14)
            d = xs * w.astype(wide)
            exact_dot = complex(math.fsum(d.real), math.fsum(d.imag)) if cplx else math.fsum(d)
            dabs = float(np.sum(np.abs(d)))
            assert _close(v.dot_product(DspVec(w)), exact_dot, dabs, tol)
            assert _close(v.dot_product(DspVec(w), prec=True), exact_dot, dabs, tol)
    # doc examples (statistics.rs:98-131)
    c = DspVec(np.array([1 + 2j, 3 + 4j, 5 + 6j], dtype=np.complex64 if dtype == np.float32 else np.complex128))
    assert c.sum() == 9 + 12j and c.sum(squared=True) == -21 + 88j
    # error behaviour (dot_products.rs:77-79,125-131)
    L = bd.lib()
    s = "32" if dtype == np.float32 else "64"
    r = DspVec(np.ones(4, dtype=dtype))
    assert getattr(L, "real_dot_product" + s)(c._h, c._h).result_code == 4
    assert getattr(L, "complex_dot_product" + s)(r._h, r._h).result_code == 3
    assert getattr(L, "complex_dot_product" + s)(c._h, r._h).result_code == 2


@pytest.mark.parametrize("dtype", DT)
def test_statistics(dtype):  # statistics.rs:45-66 (doc example), :179-440
    ct = np.complex64 if dtype == np.float32 else np.complex128
    st = DspVec(np.array([1 + 2j, 3 + 4j, 5 + 6j], dtype=ct)).statistics()
    assert st["sum"] == 9 + 12j and st["count"] == 3 and st["average"] == 3 + 4j
    assert abs(st["rms"] - (3.4027193 + 4.3102784j)) < 1e-4
    assert st["min"] == 1 + 2j and st["min_index"] == 0 and st["max"] == 5 + 6j and st["max_index"] == 2
    rng = np.random.default_rng(9)
    tol = 1e-5 if dtype == np.float32 else 1e-12
    for n in (1, 1000, 300007):
        for cplx in (False, True):
            x = rand_c(rng, n, dtype, -10, 10) if cplx else rng.uniform(-10, 10, n).astype(dtype)
            for prec in (False, True):
                got = DspVec(x).statistics(prec=prec)
                ref = o.statistics(x)
                assert got["count"] == n and got["min_index"] == ref["min_index"] and got["max_index"] == ref["max_index"]
                assert got["min"] == ref["min"] and got["max"] == ref["max"]
                scale = float(np.sum(np.abs(x.astype(np.complex128))))
                assert _close(got["sum"], ref["sum"], scale, tol)
                assert _close(got["average"], ref["average"], scale / n, tol)
                assert _close(got["rms"], ref["rms"], abs(ref["rms"]) + 1e-30, 10 * tol)
            parts = 4 if n > 1 else 1
            got = DspVec(x).statistics_split(parts)
            ref = o.statistics_split(x, parts)
            for g, r in zip(got, ref):
                assert g["count"] == r["count"] and g["min_index"] == r["min_index"] and g["max_index"] == r["max_index"]
                assert g["min"] == r["min"] and g["max"] == r["max"]
                assert _close(g["sum"], r["sum"], float(np.sum(np.abs(x.astype(np.complex128)))), tol)
            got = DspVec(x).statistics_split(parts, prec=True)
            for g, r in zip(got, ref):
                assert g["count"] == r["count"] and _close(g["sum"], r["sum"], float(np.sum(np.abs(x.astype(np.complex128)))), 1e-13)
    # doc example statistics_split (statistics.rs:82-92): 3 complex values into 2 parts
    got = DspVec(np.array([1 + 2j, 3 + 4j, 5 + 6j], dtype=ct)).statistics_split(2)
    assert got[0]["sum"] == 6 + 8j and got[1]["sum"] == 3 + 4j
    # first occurrence wins for equal extremes; empty vector -> count 0 and NaN average
    st = DspVec(np.array([1.0, 5.0, -3.0, 5.0, -3.0], dtype=dtype)).statistics()
    assert st["max_index"] == 1 and st["min_index"] == 2
    with pytest.raises(bd.DspError):
        DspVec(np.ones(64, dtype=dtype)).statistics_split(17)                           # InvalidArgumentLength
    e = DspVec(np.ones(4, dtype=dtype))
    e.set_len(0)
    st = e.statistics()
    assert st["count"] == 0 and math.isnan(st["average"]) and math.isnan(st["rms"])
