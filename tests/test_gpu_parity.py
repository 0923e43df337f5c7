"""GPU parity tests: the CUDA path (through the C ABI, via ctypes) against the CPU oracle and the
reference's golden vectors.  Tolerances are BASELINE.json's: relative L2 <= 1e-5*log2(N) for f32,
<= 1e-12*log2(N) for f64, <= 4 ulp for elementwise operations."""
import math

import numpy as np
import pytest

import basic_dsp_b200 as bd
from basic_dsp_b200 import DspVec
from oracle import dsp_oracle as o

pytestmark = pytest.mark.gpu


def tol(n, dtype):
    return (1e-5 if dtype == np.float32 else 1e-12) * max(1.0, math.log2(max(n, 2)))


def rand_c(rng, n, dtype):
    ct = np.complex64 if dtype == np.float32 else np.complex128
    return (rng.uniform(-10, 10, n) + 1j * rng.uniform(-10, 10, n)).astype(ct)


def vals(kats, key):
    return np.array(kats[key]["values"], dtype=np.float64)


def cplx(interleaved):
    a = np.asarray(interleaved, dtype=np.float64)
    return a[0::2] + 1j * a[1::2]


# --------------------------------------------------------------------------------------------------
# transforms
# --------------------------------------------------------------------------------------------------
def test_doc_vectors(kats):
    for key, fn in (("doc_plain_fft", "plain_fft"), ("doc_fft", "fft")):
        v = DspVec(cplx(kats[key]["input"]).astype(np.complex64))
        got = getattr(v, fn)().to_numpy()
        assert np.allclose(got, cplx(kats[key]["values"]), atol=1e-4), key
        assert v.domain() == bd.FREQ
    for key, fn in (("doc_plain_ifft", "plain_ifft"), ("doc_ifft", "ifft")):
        v = DspVec(cplx(kats[key]["input"]).astype(np.complex64), domain=bd.FREQ)
        got = getattr(v, fn)().to_numpy()
        assert np.allclose(got, cplx(kats[key]["values"]), atol=1e-4), key
        assert v.domain() == bd.TIME  # deliberate deviation Q2


def test_fft_vector64_golden(kats):  # tests/time_freq_test.rs:45-120
    n = np.arange(64, dtype=np.float64)
    x = np.cos(n * 0.1 * 2.0 * np.pi + 0.25)
    v = DspVec(x, dtype=np.float64).to_complex().fft().magnitude()
    assert np.max(np.abs(v.to_numpy() - vals(kats, "fft_vector64"))) < 1e-6
    # real input goes through zero_interleave inside plain_fft (time_to_freq.rs:147-150)
    v2 = DspVec(x, dtype=np.float64).fft().magnitude()
    assert np.max(np.abs(v2.to_numpy() - vals(kats, "fft_vector64"))) < 1e-6


SIZES = [1, 2, 4, 8, 16, 32, 128, 256, 512, 1024, 2048, 4096, 8192, 16384,      # single CTA
         1 << 15, 1 << 16, 1 << 17, 1 << 18, 1 << 19, 1 << 20, 1 << 21, 1 << 22,                           # multi-pass
         3, 5, 6, 12, 24, 3 * 64, 5 * 1024, 7 * 4096, 3 * (1 << 16), 15 * (1 << 14),  # q * 2^k
         17 * 29, 1001, 9973, 10007 * 3, 127 * 127, 1000003]                    # Bluestein


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n", SIZES)
def test_plain_fft_sizes(n, dtype):
    rng = np.random.default_rng(n)
    x = rand_c(rng, n, dtype)
    v = DspVec(x, delta=0.5)
    got = v.plain_fft().to_numpy()
    assert o.rel_l2(got, o.plain_fft(x)) <= tol(n, dtype)
    assert v.domain() == bd.FREQ and v.is_complex()
    T = dtype
    assert v.delta() == float(o.delta_after_fft(0.5, n, T))
    # inverse (unnormalised) brings back n * x
    back = v.plain_ifft().to_numpy()
    assert o.rel_l2(back / n, x) <= 2 * tol(n, dtype)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n", [7, 8, 64, 1000, 1001, 4096, 1 << 15, 65536, 3 * 4096, 1 << 17, 1 << 18, 1 << 19, 1 << 20])
def test_fft_ifft_shifted(n, dtype):
    rng = np.random.default_rng(100 + n)
    x = rand_c(rng, n, dtype)
    v = DspVec(x)
    X = v.fft().to_numpy()
    assert o.rel_l2(X, o.fft(x)) <= tol(n, dtype)
    y = v.ifft().to_numpy()
    assert o.rel_l2(y, x) <= 2 * tol(n, dtype)
    assert o.rel_l2(DspVec(X.copy(), domain=bd.FREQ).ifft().to_numpy(), o.ifft(X)) <= tol(n, dtype)


@pytest.mark.parametrize("n", [64, 1001, 4096, 1 << 15, 1 << 16, 1 << 18, 1 << 20, 1 << 21])
def test_real_input_fft(n):  # tests/real_test.rs:581-605
    rng = np.random.default_rng(n)
    x = rng.uniform(-10, 10, n).astype(np.float32)
    got = DspVec(x).plain_fft().to_numpy()
    assert len(got) == n
    assert o.rel_l2(got, o.plain_fft(x)) <= tol(n, np.float32)


@pytest.mark.parametrize("n,rows", [(64, 2048), (128, 1024), (256, 256), (512, 128), (1024, 256), (2048, 32), (4096, 64), (8192, 8), (16384, 4), (1 << 15, 4), (1 << 16, 4), (1 << 19, 2), (1 << 20, 2), (1 << 22, 1), (512, 7)])
def test_real_input_rows(n, rows):
    """rows of real scalars (BDSP_F_REAL_INPUT): complexifying pass + packed passes in the throughput regime."""
    rng = np.random.default_rng(n + rows)
    L = bd.lib()
    x = rng.uniform(-10, 10, n * rows).astype(np.float32)
    v = DspVec(x)
    out = DspVec.zeros(2 * n * rows, is_complex=True, dtype=np.float32)
    pin, pout = v._fn("bdsp_device_ptr")(v._h), out._fn("bdsp_device_ptr")(out._h)
    assert L.bdsp_fft_rows_c32(pin, pout, n, rows, bd.F_REAL_INPUT) == 0
    ref = np.fft.fft(x.reshape(rows, n).astype(np.float64), axis=1)
    assert o.rel_l2(out.to_numpy().reshape(rows, n), ref) <= tol(n, np.float32)
    assert L.bdsp_fft_rows_c32(pin, pout, n, rows, bd.F_REAL_INPUT | bd.F_SHIFT) == 0
    assert o.rel_l2(out.to_numpy().reshape(rows, n), np.fft.fftshift(ref, axes=1)) <= tol(n, np.float32)
    mag = DspVec.zeros(n * rows, dtype=np.float32)
    assert L.bdsp_fft_rows_c32(pin, mag._fn("bdsp_device_ptr")(mag._h), n, rows, bd.F_REAL_INPUT | bd.F_SHIFT | bd.F_MAGNITUDE) == 0
    assert o.rel_l2(mag.to_numpy().reshape(rows, n), np.abs(np.fft.fftshift(ref, axes=1))) <= tol(n, np.float32)


def test_fft_wrong_domain_marks_invalid():
    v = DspVec(np.ones(8, dtype=np.complex64), domain=bd.FREQ)
    assert v.result_code_of("fft") == -1
    assert v.len() == 0 and math.isnan(v.delta())
    v = DspVec(np.ones(8, dtype=np.complex64), domain=bd.TIME)
    assert v.result_code_of("ifft") == -1


@pytest.mark.parametrize("n", [2, 5, 8, 9, 1000, 1001])
def test_swap_halves(n):
    x = np.arange(n, dtype=np.float32)
    assert np.array_equal(DspVec(x).swap_halves().to_numpy(), np.fft.fftshift(x))
    c = (np.arange(n) + 1j * np.arange(n)[::-1]).astype(np.complex64)
    assert np.array_equal(DspVec(c, domain=bd.FREQ).fft_shift().to_numpy(), np.fft.fftshift(c))
    assert np.array_equal(DspVec(c, domain=bd.FREQ).ifft_shift().to_numpy(), np.fft.ifftshift(c))


def test_fft_rows_batched():
    rng = np.random.default_rng(5)
    L = bd.lib()
    for n, rows in [(16384, 8), (1024, 33), (64, 100), (1 << 15, 3), (1 << 16, 5), (1 << 17, 3), (1 << 18, 5), (1 << 19, 3), (1 << 20, 3),
                    (64, 128), (128, 96), (256, 48), (512, 64), (1024, 32), (2048, 6)]:
        x = rand_c(rng, n * rows, np.float32)
        v = DspVec(x)
        out = DspVec.zeros(n * rows, dtype=np.float32)
        rc = L.bdsp_fft_rows_c32(v._fn("bdsp_device_ptr")(v._h), out._fn("bdsp_device_ptr")(out._h), n, rows,
                                 bd.F_SHIFT | bd.F_MAGNITUDE)
        assert rc == 0
        got = out.to_numpy().reshape(rows, n)
        ref = np.abs(np.fft.fftshift(np.fft.fft(x.reshape(rows, n).astype(np.complex128), axis=1), axes=1))
        assert o.rel_l2(got, ref) <= tol(n, np.float32)


@pytest.mark.parametrize("n", [16384, 1 << 15, 1 << 16, 1 << 17, 1 << 18, 1 << 19, 1 << 20])
def test_fft_magnitude_fused(n):
    rng = np.random.default_rng(6)
    x = rand_c(rng, n, np.float32)
    got = DspVec(x).fft_magnitude()
    assert not got.is_complex() and got.len() == n
    assert o.rel_l2(got.to_numpy(), np.abs(o.fft(x))) <= tol(n, np.float32)


@pytest.mark.parametrize("n", [64, 128, 256, 512, 1024, 2048, 4096, 8192])
def test_small_row_batches_all_flag_combinations(n):
    """several rows per CTA (fftp.cu NATQ modes) and the 4096 / 8192-point kernels: forward, shifted, inverse, ifft."""
    rng = np.random.default_rng(n + 1)
    L = bd.lib()
    rows = 3 * (4096 // n if n < 4096 else 1) * 8
    x = rand_c(rng, n * rows, np.float32)
    v = DspVec(x)
    out = DspVec.zeros(2 * n * rows, is_complex=True, dtype=np.float32)
    pin, pout = v._fn("bdsp_device_ptr")(v._h), out._fn("bdsp_device_ptr")(out._h)
    xr = x.reshape(rows, n).astype(np.complex128)
    cases = [(0, np.fft.fft(xr, axis=1)),
             (bd.F_SHIFT, np.fft.fftshift(np.fft.fft(xr, axis=1), axes=1)),
             (bd.F_INVERSE, np.fft.ifft(xr, axis=1) * n),
             (bd.F_INVERSE | bd.F_SHIFT, np.fft.ifft(np.fft.ifftshift(xr, axes=1), axis=1))]
    for flags, ref in cases:
        assert L.bdsp_fft_rows_c32(pin, pout, n, rows, flags) == 0
        assert o.rel_l2(out.to_numpy().reshape(rows, n), ref) <= tol(n, np.float32), (n, flags)


@pytest.mark.parametrize("n", [1 << 15, 1 << 16, 1 << 17, 1 << 18, 1 << 19, 1 << 20])
def test_two_pass_rows_plain_and_inverse(n):
    """packed two-pass path (fftp.cu): plain forward / inverse over several rows, unshifted."""
    rng = np.random.default_rng(n)
    L = bd.lib()
    rows = 3
    x = rand_c(rng, n * rows, np.float32)
    v = DspVec(x)
    out = DspVec.zeros(2 * n * rows, is_complex=True, dtype=np.float32)
    pin, pout = v._fn("bdsp_device_ptr")(v._h), out._fn("bdsp_device_ptr")(out._h)
    assert L.bdsp_fft_rows_c32(pin, pout, n, rows, 0) == 0
    ref = np.fft.fft(x.reshape(rows, n).astype(np.complex128), axis=1)
    assert o.rel_l2(out.to_numpy().reshape(rows, n), ref) <= tol(n, np.float32)
    assert L.bdsp_fft_rows_c32(pin, pout, n, rows, bd.F_INVERSE) == 0
    ref = np.fft.ifft(x.reshape(rows, n).astype(np.complex128), axis=1) * n
    assert o.rel_l2(out.to_numpy().reshape(rows, n), ref) <= tol(n, np.float32)


# --------------------------------------------------------------------------------------------------
# convolution
# --------------------------------------------------------------------------------------------------
def test_convolve_signal_kats(kats):
    a = np.arange(10, dtype=np.float32).astype(np.complex64)
    b = np.zeros(10, dtype=np.complex64); b[4] = 1
    got = DspVec(a).convolve_signal(DspVec(b)).magnitude().to_numpy()
    assert np.allclose(got, vals(kats, "shift_left_by_1_as_conv"), atol=1e-4)
    b = np.array([0, 0, 1], dtype=np.complex64)
    got = DspVec(a).convolve_signal(DspVec(b)).magnitude().to_numpy()
    assert np.allclose(got, vals(kats, "shift_left_by_1_as_conv_shorter"), atol=1e-4)
    # convolution.rs:737-775
    n = 11
    x = np.zeros(n, dtype=np.complex64); x[n // 2] = 1
    h = np.array([o.sinc_impulse(np.float32(v * 0.5), np.float32) for v in np.arange(-5.0, 6.0)]).astype(np.complex64)
    got = DspVec(x).convolve_signal(DspVec(h)).magnitude().to_numpy()
    assert np.allclose(got, vals(kats, "convolve_complex_vectors32"), atol=1e-4)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n,l", [(64, 1), (64, 2), (65, 7), (300, 24), (300, 25), (1000, 129), (5000, 1023),
                                 (9500, 100), (20000, 1023), (4096, 4096), (10000, 2500), (40000, 5001),
                                 (3 * 4096 + 5, 2000), (20000, 2046), (20000, 2047), (9000, 1500), (4096, 1024)])
def test_convolve_signal_vs_oracle(n, l, dtype):
    rng = np.random.default_rng(n * 7 + l)
    x = rand_c(rng, n, dtype)
    h = (rand_c(rng, l, dtype) / 10).astype(x.dtype)
    got = DspVec(x).convolve_signal(DspVec(h)).to_numpy()
    ref = o.convolve_signal_direct(x, h) if n * l < 3e7 else o.convolve_signal(x, h)
    assert o.rel_l2(got, ref) <= tol(max(n, 4096), dtype)


@pytest.mark.parametrize("n,l", [(100, 6), (5000, 63), (20000, 200)])
def test_convolve_signal_real_vectors(n, l):
    rng = np.random.default_rng(n + l)
    x = rng.uniform(-10, 10, n).astype(np.float32)
    h = rng.uniform(-1, 1, l).astype(np.float32)
    got = DspVec(x).convolve_signal(DspVec(h)).to_numpy()
    assert o.rel_l2(got, o.convolve_signal_direct(x, h)) <= tol(4096, np.float32)


def test_convolve_signal_c2a_full_size():
    """BASELINE config C2a: 2^20 c32 with the 1023-tap raised cosine h[k] = RC_0.35((k-511)*0.25)."""
    n, l = 1 << 20, 1023
    rng = np.random.default_rng(20260102)
    x = rand_c(rng, n, np.float32)
    h = np.array([o.raised_cosine_impulse(np.float32((k - 511) * 0.25), 0.35, np.float32) for k in range(l)])
    hv = DspVec(h.astype(np.complex64))
    v = DspVec(x)
    got = v.convolve_signal(hv).to_numpy()
    ref = o.convolve_signal(x, h.astype(np.complex128))
    assert o.rel_l2(got, ref) <= tol(n, np.float32)
    # size-independent properties: linearity and shift equivariance of the circular convolution
    x2 = rand_c(rng, n, np.float32)
    y2 = DspVec(x2).convolve_signal(hv).to_numpy()
    ys = DspVec((x + 2 * x2).astype(np.complex64)).convolve_signal(hv).to_numpy()
    assert o.rel_l2(ys, got.astype(np.complex128) + 2 * y2.astype(np.complex128)) <= 4 * tol(n, np.float32)
    yr = DspVec(np.roll(x, 12345)).convolve_signal(hv).to_numpy()
    assert o.rel_l2(yr, np.roll(got, 12345)) <= 4 * tol(n, np.float32)
    # second call reuses the cached impulse-response spectrum
    again = DspVec(x).convolve_signal(hv).to_numpy()
    assert np.array_equal(again, got)


def test_convolve_signal_rows_api():
    L = bd.lib()
    n, rows, l = 20000, 5, 1023
    rng = np.random.default_rng(9)
    x = rand_c(rng, n * rows, np.float32)
    h = (rand_c(rng, l, np.float32) / 10).astype(np.complex64)
    xv, hv, out = DspVec(x), DspVec(h), DspVec.zeros(2 * n * rows, is_complex=True)
    dp = lambda v: v._fn("bdsp_device_ptr")(v._h)
    plan = L.bdsp_conv_plan_create_c32(dp(hv), l)
    assert plan
    assert L.bdsp_convolve_signal_rows_c32(dp(xv), dp(out), n, rows, plan) == 0
    got = out.to_numpy().reshape(rows, n)
    for r in range(rows):
        assert o.rel_l2(got[r], o.convolve_signal(x.reshape(rows, n)[r], h)) <= tol(4096, np.float32)
    L.bdsp_conv_plan_destroy(plan)


def test_convolve_signal_errors():
    x = DspVec(np.ones(10, dtype=np.complex64))
    assert x.result_code_of("convolve_signal", DspVec(np.ones(11, dtype=np.complex64))) == o.ERR_INVALID_ARG_LEN
    assert x.result_code_of("convolve_signal", DspVec(np.ones(5, dtype=np.float32))) == o.ERR_META_DATA
    assert x.result_code_of("convolve_signal", DspVec(np.ones(5, dtype=np.complex64), delta=2.0)) == o.ERR_META_DATA
    f = DspVec(np.ones(10, dtype=np.complex64), domain=bd.FREQ)
    assert f.result_code_of("convolve_signal", DspVec(np.ones(5, dtype=np.complex64), domain=bd.FREQ)) == o.ERR_MUST_BE_TIME
    assert x.result_code_of("convolve_signal", DspVec(np.ones(5, dtype=np.complex64))) == 0


def test_convolve_function_kats(kats):
    x = np.zeros(10, dtype=np.float32); x[5] = 1
    got = DspVec(x).convolve(bd.RAISED_COSINE, 0.35, 0.2, 5).to_numpy()
    assert np.max(np.abs(got - vals(kats, "convolve_real_time_and_time32"))) < 1e-4
    n = 11
    c = np.zeros(n, dtype=np.complex64); c[n // 2] = 1
    got = DspVec(c).convolve(bd.SINC, 0.0, 0.5, n // 2).magnitude().to_numpy()
    assert np.max(np.abs(got - vals(kats, "convolve_complex_time_and_time32"))) < 1e-4


@pytest.mark.parametrize("n,ratio,length,cplx_", [(1 << 20, 0.25, 31, True), (5000, 1.0, 12, True), (3000, 1.0, 12, False),
                                                  (1500, 0.5, 40, False), (20, 0.5, 200, True)])
def test_convolve_function_vs_oracle(n, ratio, length, cplx_):
    """(2^20, 0.25, 31) is BASELINE config C2b: 63-tap raised-cosine `convolve`."""
    rng = np.random.default_rng(n + length)
    x = rand_c(rng, n, np.float32) if cplx_ else rng.uniform(-10, 10, n).astype(np.float32)
    got = DspVec(x).convolve(bd.RAISED_COSINE, 0.35, ratio, length).to_numpy()
    ref = o.convolve_function(x, lambda t: o.raised_cosine_impulse(t, 0.35, np.float32), ratio, length, np.float32)
    assert o.rel_l2(got, ref) <= tol(4096, np.float32)
    cb = DspVec(x).convolve(lambda t: float(o.raised_cosine_impulse(t, 0.35, np.float32)), 0.0, ratio, length).to_numpy()
    assert o.rel_l2(cb, ref) <= tol(4096, np.float32)


def test_convolve_complex_callback():
    rng = np.random.default_rng(77)
    x = rand_c(rng, 3000, np.float32)
    fn = lambda t: complex(o.sinc_impulse(t, np.float32), 0.5 * o.sinc_impulse(t * 0.5, np.float32))
    got = DspVec(x).convolve_complex(fn, 0.5, 10).to_numpy()
    idx = np.arange(3000)
    ref = np.zeros(3000, dtype=np.complex128)
    for m in range(-10, 11):
        ref += x[(idx + m) % 3000].astype(np.complex128) * fn(np.float32(-m * 0.5))
    assert o.rel_l2(got, ref) <= tol(4096, np.float32)


def test_multiply_frequency_response(kats):
    for n, key in ((5, "convolve_complex_freq_and_freq32"), (6, "convolve_complex_freq_and_freq_even32")):
        v = DspVec(np.ones(n, dtype=np.complex64) * (1 + 1j), domain=bd.FREQ)
        got = v.multiply_frequency_response(bd.RAISED_COSINE, 1.0, 2.0).to_numpy()
        flat = np.empty(2 * n); flat[0::2], flat[1::2] = got.real, got.imag
        assert np.max(np.abs(flat - vals(kats, key))) < 1e-4
    rng = np.random.default_rng(3)
    X = rand_c(rng, 1001, np.float32)
    got = DspVec(X, domain=bd.FREQ).multiply_frequency_response(bd.RAISED_COSINE, 0.35, 0.8).to_numpy()
    ref = o.multiply_frequency_response(X, lambda t: o.raised_cosine_freq(t, 0.35, np.float32), 0.8, np.float32)
    assert o.rel_l2(got, ref) <= 1e-6
    cb = DspVec(X, domain=bd.FREQ).multiply_frequency_response(lambda t: float(o.raised_cosine_freq(t, 0.35, np.float32)), 0.0, 0.8).to_numpy()
    assert o.rel_l2(cb, ref) <= 1e-6
    assert DspVec(X).result_code_of("multiply_frequency_response", 0, 0.0, 1.0) == -1


# --------------------------------------------------------------------------------------------------
# interpolation
# --------------------------------------------------------------------------------------------------
def test_interpolatef_kats(kats):
    for n, key in ((6, "interpolatef_by_integer_sinc_even_test"), (7, "interpolatef_by_integer_sinc_odd_test")):
        x = np.zeros(n, dtype=np.complex64); x[n // 2] = 1
        got = DspVec(x).interpolatef(bd.SINC, 0.0, 2.0, 0.0, n).to_real().to_numpy()
        assert np.max(np.abs(got - vals(kats, key))) < 0.1
    x = np.zeros(6, dtype=np.complex64); x[3] = 1
    got = DspVec(x).interpolatef(bd.SINC, 0.0, 13.0 / 6.0, 0.0, 6).to_real().to_numpy()
    e = vals(kats, "interpolatef_by_fractional_sinc_test")
    assert len(got) == len(e) and np.max(np.abs(got - e)) < 0.1
    got = DspVec(x).interpolatef(bd.SINC, 0.0, 2.0, 1.0, 6).magnitude().to_numpy()
    assert np.max(np.abs(got - vals(kats, "interpolatef_delayed_sinc_test"))) < 0.1


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n,F,L,delay,cplx_", [(1000, 4, 12, 0.0, False), (1000, 4, 12, 0.0, True), (777, 3, 5, 0.3, False),
                                               (2000, 2, 10, 0.0, True), (600, 8, 20, 0.0, False), (500, 16, 4, 0.0, False),
                                               (1 << 18, 4, 12, 0.0, False), (301, 7, 13, -0.2, True)])
def test_interpolatef_integer_vs_oracle(n, F, L, delay, cplx_, dtype):
    rng = np.random.default_rng(n + F)
    x = rand_c(rng, n, dtype) if cplx_ else rng.uniform(-10, 10, n).astype(dtype)
    got = DspVec(x).interpolatef(bd.SINC, 0.0, float(F), delay, L).to_numpy()
    ref = o.interpolatef(x, lambda t: o.sinc_impulse(t, dtype), float(F), delay, L, dtype)
    assert len(got) == len(ref)
    assert o.rel_l2(got, ref) <= tol(4096, dtype)
    # 4-argument custom callback path builds the same tables on the host
    if n <= 2000:
        cb = DspVec(x).interpolatef(lambda t: float(o.sinc_impulse(t, dtype)), 0.0, float(F), delay, L).to_numpy()
        assert o.rel_l2(cb, ref) <= tol(4096, dtype)


@pytest.mark.parametrize("n,factor,L,delay,kind", [(400, 2.5, 8, 0.0, 0), (300, 13.0 / 6.0, 6, 0.0, 0), (500, 0.5, 10, 0.0, 1),
                                                   (100, 3.0, 10, 0.25, 1)])
def test_interpolatef_scalar_path_vs_oracle(n, factor, L, delay, kind):
    rng = np.random.default_rng(n)
    x = rand_c(rng, n, np.float32)
    f = (lambda t: o.sinc_impulse(t, np.float32)) if kind == 0 else (lambda t: o.raised_cosine_impulse(t, 0.35, np.float32))
    got = DspVec(x).interpolatef(kind, 0.35, np.float32(factor), delay, L).to_numpy()
    ref = o.interpolatef(x, f, np.float32(factor), delay, L, np.float32)
    assert len(got) == len(ref)
    assert o.rel_l2(got, ref) <= 1e-4  # taps evaluated with device sin/cos in f32


def test_interpolate_lin(kats):
    x = np.array([-1.0, -2.0, -1.0, 0.0, 1.0, 3.0, 4.0], dtype=np.float32)
    got = DspVec(x).interpolate_lin(4.0, 0.0).to_numpy()
    assert np.max(np.abs(got - vals(kats, "linear_test"))) < 1e-6
    rng = np.random.default_rng(1)
    for dtype in (np.float32, np.float64):
        for n, F, d in [(1000, 4.0, 0.0), (12345, 3.0, 0.0), (999, 2.5, 0.25), (1 << 20, 4.0, 0.0)]:
            x = rng.uniform(-10, 10, n).astype(dtype)
            got = DspVec(x).interpolate_lin(F, d).to_numpy()
            ref = o.interpolate_lin(x, F, d, dtype)
            assert len(got) == len(ref)
            assert np.array_equal(got, ref), (dtype, n, F, d)   # bit exact: same IEEE operations
    assert DspVec(np.ones(8, dtype=np.complex64)).result_code_of("interpolate_lin", 2.0, 0.0) == -1


def test_interpolate_lin_c4b_full_size_counter_saturation():
    """BASELINE config C4b (2^24 real f32, x4): beyond output 2^24 the reference's f32 counter is stuck (Q7)."""
    n = 1 << 24
    rng = np.random.default_rng(4)
    x = rng.uniform(-10, 10, n).astype(np.float32)
    got = DspVec(x).interpolate_lin(4.0, 0.0).to_numpy()
    ref = o.interpolate_lin(x, 4.0, 0.0, np.float32)
    assert len(got) == len(ref) == 4 * (n - 1) + 1
    assert np.array_equal(got, ref)
    assert np.all(got[(1 << 24):-1] == got[1 << 24])


# --------------------------------------------------------------------------------------------------
# elementwise chain (<= 4 ulp)
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_elementwise_ulp(dtype):
    rng = np.random.default_rng(42)
    n = 100003
    x, w = rand_c(rng, n, dtype), rand_c(rng, n, dtype)
    assert np.array_equal(DspVec(x).scale(2.5).to_numpy(), o.real_scale(x, 2.5, dtype))
    assert np.array_equal(DspVec(x).scale(complex(1.5, -0.25)).to_numpy(), o.complex_scale(x, complex(1.5, -0.25), dtype))
    assert np.array_equal(DspVec(x).mul(DspVec(w)).to_numpy(), o.mul(x, w, dtype))
    assert np.array_equal(DspVec(x).add(DspVec(w)).to_numpy(), o.add(x, w, dtype))
    assert np.array_equal(DspVec(x).sub(DspVec(w)).to_numpy(), o.sub(x, w, dtype))
    d = DspVec(x).div(DspVec(w)).to_numpy()
    dr = o.div(x, w, dtype)
    assert o.ulp_diff(d.real, dr.real, dtype).max() <= 4 and o.ulp_diff(d.imag, dr.imag, dtype).max() <= 4
    assert o.ulp_diff(DspVec(x).magnitude().to_numpy(), o.magnitude(x, dtype), dtype).max() <= 4
    assert np.array_equal(DspVec(x).magnitude_squared().to_numpy(), o.magnitude_squared(x, dtype))
    assert o.ulp_diff(DspVec(x).phase().to_numpy(), o.phase(x, dtype), dtype).max() <= 4
    assert np.array_equal(DspVec(x).to_real().to_numpy(), x.real)
    assert np.array_equal(DspVec(x).to_imag().to_numpy(), x.imag)
    assert np.array_equal(DspVec(x).conj().to_numpy(), np.conj(x))
    r = rng.uniform(-10, 10, n).astype(dtype)
    assert np.array_equal(DspVec(r).mul(DspVec(r[::-1].copy())).to_numpy(), o.mul(r, r[::-1], dtype))
    assert np.array_equal(DspVec(r).offset(1.25).to_numpy(), (r + dtype(1.25)).astype(dtype))
    assert np.array_equal(DspVec(r).to_complex().to_numpy(), r.astype(x.dtype))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_fused_chain_equals_sequential_calls(dtype):
    rng = np.random.default_rng(8)
    n = 70001
    x, w = rand_c(rng, n, dtype), rand_c(rng, n, dtype)
    c = complex(0.75, -1.5)
    seq = DspVec(x).scale(c).mul(DspVec(w))
    m1, p1 = DspVec.zeros(0, dtype=dtype), DspVec.zeros(0, dtype=dtype)
    assert seq.get_mag_phase(m1, p1) == o.CONVERT_VOID_OK
    m2, p2 = DspVec.zeros(0, dtype=dtype), DspVec.zeros(0, dtype=dtype)
    v = DspVec(x)
    v.scale_mul_mag_phase(c, DspVec(w), m2, p2, write_back=True)
    assert np.array_equal(m1.to_numpy(), m2.to_numpy()) and np.array_equal(p1.to_numpy(), p2.to_numpy())
    assert np.array_equal(v.to_numpy(), seq.to_numpy())
    ref = o.mul(o.complex_scale(x, c, dtype), w, dtype)
    assert o.ulp_diff(m2.to_numpy(), o.magnitude(ref, dtype), dtype).max() <= 4
    assert o.ulp_diff(p2.to_numpy(), o.phase(ref, dtype), dtype).max() <= 4


def test_elementwise_errors_and_getters():
    a = DspVec(np.ones(8, dtype=np.complex64))
    assert a.result_code_of("mul", DspVec(np.ones(7, dtype=np.complex64))) == o.ERR_SAME_SIZE
    assert a.result_code_of("mul", DspVec(np.ones(16, dtype=np.float32))) == o.ERR_META_DATA
    assert a.result_code_of("add", DspVec(np.ones(8, dtype=np.complex64), domain=bd.FREQ)) == o.ERR_META_DATA
    r = DspVec(np.ones(8, dtype=np.float32))
    assert r.result_code_of("complex_scale", 1.0, 1.0) == -1          # assert_complex! marks the vector invalid
    assert r.len() == 0 and math.isnan(r.delta())
    src = DspVec(np.array([3 - 4j, -3 + 4j], dtype=np.complex64), delta=0.25)
    dst = DspVec.zeros(0)
    assert src.get_magnitude(dst) == o.CONVERT_VOID_OK               # Q8
    assert np.array_equal(dst.to_numpy(), [5.0, 5.0]) and dst.delta() == 0.25
    cdst = DspVec.zeros(4, is_complex=True)
    assert src.get_phase(cdst) == o.CONVERT_VOID_OK and cdst.len() == 0


def test_metadata_and_host_access():
    v = DspVec.zeros(10, init=1.5)
    assert v.len() == 10 and v.points() == 10 and not v.is_complex() and v.domain() == bd.TIME
    assert v.get_value(3) == 1.5
    v.set_value(3, 2.0)
    assert np.array_equal(v.data_via_host_mirror(), [1.5, 1.5, 1.5, 2.0] + [1.5] * 6)
    c = DspVec.zeros(7, is_complex=True)          # odd length complex -> valid_len 0 (support_std.rs:369)
    assert c.len() == 0 and c.alloc_len() >= 7
    c = DspVec.zeros(8, is_complex=True)
    c.set_len(5)                                    # odd -> ignored (InputMustHaveAnEvenLength)
    assert c.len() == 8
    c.set_len(20)                                   # grows, zero filled
    assert c.len() == 20 and np.array_equal(c.to_numpy(), np.zeros(10))
    data = np.arange(5, dtype=np.float32)
    L = bd.lib()
    import ctypes
    res = L.overwrite_data32(v._h, data.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), 5)
    assert res.result_code == 0 and np.array_equal(v.to_numpy()[:5], data)
    big = np.arange(10, dtype=np.float32)
    res = L.overwrite_data32(v._h, big.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), 10)
    assert res.result_code == o.ERR_INVALID_ARG_LEN  # Q9: strict len < vec.len()
    w = v.clone()
    assert np.array_equal(w.to_numpy(), v.to_numpy())
    assert bd.kernel_launch_count() > 0


def test_zero_pad_and_interleave():
    x = np.arange(1, 6, dtype=np.float32)
    assert np.array_equal(DspVec(x).zero_pad(8, 0).to_numpy(), [1, 2, 3, 4, 5, 0, 0, 0])
    assert np.array_equal(DspVec(x).zero_pad(9, 1).to_numpy(), [0, 0, 1, 2, 3, 4, 5, 0, 0])
    assert np.array_equal(DspVec(x).zero_interleave(3).to_numpy(), [1, 0, 0, 2, 0, 0, 3, 0, 0, 4, 0, 0, 5, 0, 0])
    c = np.array([1 + 2j, 3 + 4j], dtype=np.complex64)
    assert np.array_equal(DspVec(c).zero_interleave(2).to_numpy(), [1 + 2j, 0, 3 + 4j, 0])
    assert DspVec(x).result_code_of("zero_pad", 5, 0) == o.ERR_INVALID_ARG_LEN


# --------------------------------------------------------------------------------------------------
# next rows of the scope table: windows, correlation, reverse, decimatei
# --------------------------------------------------------------------------------------------------
def test_windowed_fft_golden(kats):  # tests/time_freq_test.rs:122-197
    n = np.arange(64, dtype=np.float64)
    x = np.cos(n * 0.1 * 2.0 * np.pi + 0.25)
    got = DspVec(x, dtype=np.float64).to_complex().windowed_fft(1).magnitude().to_numpy()
    assert np.max(np.abs(got - vals(kats, "windowed_fft_vector64"))) < 1e-6


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("kind", [0, 1, 2, 3])
def test_apply_window_and_roundtrip(kind, dtype):
    rng = np.random.default_rng(kind)
    for n in (5, 64, 1001):
        x = rand_c(rng, n, dtype)
        got = DspVec(x).apply_window(kind).to_numpy()
        ref = o.apply_window(x, kind, dtype)
        # the window values are evaluated on the host with libm's cos in precision T (as the reference does);
        # NumPy's float32 cos can differ by 1 ulp, which alpha - beta*cos(..) amplifies near the window edges
        # (BlackmanHarris edge values ~6e-5 are the result of a 4-term cancellation), so the cosine windows are
        # compared in the L2 sense and only the cos-free windows to 4 ulp
        assert o.rel_l2(got, ref) <= (1e-6 if dtype == np.float32 else 1e-14)
        r = rng.uniform(-10, 10, n).astype(dtype)
        gr, rr = DspVec(r).apply_window(kind).to_numpy(), o.apply_window(r, kind, dtype)
        assert o.rel_l2(gr, rr) <= (1e-6 if dtype == np.float32 else 1e-14)
        if kind in (0, 3):
            assert o.ulp_diff(got.real, ref.real, dtype).max() <= 4 and o.ulp_diff(got.imag, ref.imag, dtype).max() <= 4
            assert o.ulp_diff(gr, rr, dtype).max() <= 4
    x = rand_c(rng, 4096, dtype)
    if kind != 0:   # the triangular window is fine too, but keep away from tiny edge values in f32
        back = DspVec(x).windowed_fft(kind).windowed_ifft(kind).to_numpy()
        assert o.rel_l2(back, x) <= 50 * tol(4096, dtype)
    X = DspVec(x).windowed_fft(kind).to_numpy()
    assert o.rel_l2(X, o.windowed_fft(x, kind, dtype)) <= tol(4096, dtype)


def test_correlation_golden(kats):  # correlation.rs:170-215
    a = cplx(kats["time_correlation_test_a"]["values"]).astype(np.complex64)
    b = cplx(kats["time_correlation_test_b"]["values"]).astype(np.complex64)
    c = vals(kats, "time_correlation_test_c")
    prepared = DspVec(b).prepare_argument_padded()
    assert prepared.domain() == bd.FREQ and prepared.points() == 2 * len(b) - 1
    got = DspVec(a).correlate(prepared).to_numpy()
    flat = np.empty(2 * len(got)); flat[0::2], flat[1::2] = got.real, got.imag
    assert len(flat) == len(c) and np.max(np.abs(flat - c)) < 0.1
    a2 = np.array([1 + 1j, 2 + 1j, 3 + 1j], dtype=np.complex64)
    b2 = np.array([4 + 1j, 5 + 1j, 6 + 1j], dtype=np.complex64)
    got = DspVec(a2).correlate(DspVec(b2).prepare_argument_padded()).to_numpy()
    flat = np.empty(2 * len(got)); flat[0::2], flat[1::2] = got.real, got.imag
    assert np.max(np.abs(flat - vals(kats, "time_correlation_test2_c"))) < 1e-3


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n", [17, 1000, 5000, 40000])
def test_correlate_vs_oracle(n, dtype):
    rng = np.random.default_rng(n)
    a, b = rand_c(rng, n, dtype), rand_c(rng, n, dtype)
    prepared = DspVec(b).prepare_argument_padded()
    assert o.rel_l2(prepared.to_numpy(), o.prepare_argument_padded(b)) <= tol(2 * n, dtype)
    got = DspVec(a, delta=0.5).correlate(prepared)
    assert got.domain() == bd.TIME and got.delta() == 0.5
    assert o.rel_l2(got.to_numpy(), o.correlate(a, o.prepare_argument_padded(b))) <= 2 * tol(2 * n, dtype)
    p2 = DspVec(b).prepare_argument()
    assert o.rel_l2(p2.to_numpy(), o.prepare_argument(b)) <= tol(n, dtype)
    # same number of points: zero_pad_b refuses (InvalidArgumentLength), as in the reference
    assert DspVec(a).result_code_of("correlate", p2) == o.ERR_INVALID_ARG_LEN
    assert DspVec(a).result_code_of("correlate", DspVec(b)) == o.ERR_MUST_BE_TIME   # other must be a frequency vector


def test_reverse_and_decimatei():
    x = np.arange(11, dtype=np.float32)
    assert np.array_equal(DspVec(x).reverse().to_numpy(), x[::-1])
    c = (np.arange(7) + 1j * np.arange(7)[::-1]).astype(np.complex64)
    assert np.array_equal(DspVec(c).reverse().to_numpy(), c[::-1])
    assert np.array_equal(DspVec(x).decimatei(3, 1).to_numpy(), o.decimatei(x, 3, 1))
    assert np.array_equal(DspVec(c).decimatei(2, 0).to_numpy(), o.decimatei(c, 2, 0))
    big = np.arange(100003, dtype=np.float64)
    assert np.array_equal(DspVec(big).decimatei(7, 5).to_numpy(), big[5::7])


def test_interpolatei(kats):  # interpolation.rs:653-680, 722-750 + tests/interpolation_test.rs pattern
    n = 6
    x = np.zeros(n, dtype=np.complex64); x[n // 2] = 1
    got = DspVec(x).interpolatei(bd.SINC, 0.0, 2).magnitude().to_numpy()
    assert np.max(np.abs(got - vals(kats, "interpolatei_sinc_test"))) < 1e-4
    got = DspVec(x).interpolatei(bd.RAISED_COSINE, 0.4, 2).magnitude().to_numpy()
    assert np.max(np.abs(got - vals(kats, "interpolatei_rc_test"))) < 1e-4
    rng = np.random.default_rng(12)
    for dtype in (np.float32, np.float64):
        for n, F, cplx_ in [(1000, 4, True), (777, 3, True), (512, 2, False), (301, 5, False)]:
            x = rand_c(rng, n, dtype) if cplx_ else rng.uniform(-10, 10, n).astype(dtype)
            f = lambda t: o.raised_cosine_freq(t, 0.35, dtype)
            got = DspVec(x).interpolatei(bd.RAISED_COSINE, 0.35, F)
            ref = o.interpolatei(x, f, F, dtype)
            assert got.is_complex() == cplx_ and got.points() == n * F
            assert o.rel_l2(got.to_numpy(), ref) <= 4 * tol(n * F, dtype)
            cb = DspVec(x).interpolatei(lambda t: float(o.raised_cosine_freq(t, 0.35, dtype)), 0.0, F).to_numpy()
            assert o.rel_l2(cb, ref) <= 4 * tol(n * F, dtype)
    # decimatei is the inverse for a band-limited signal when the response is flat in band
    x = np.cos(2 * np.pi * 5 * np.arange(256) / 256).astype(np.float32)
    back = DspVec(x).interpolatei(bd.SINC, 0.0, 4).decimatei(4, 0).to_numpy()
    assert np.max(np.abs(back - x)) < 1e-3
    assert DspVec(x).result_code_of("interpolatei", 0, 0.0, 1) == 0   # factor <= 1: no-op


# --------------------------------------------------------------------------------------------------
# FFT-based resampling (interpolate / interpft), symmetric transforms, mirror, complex exponential
# --------------------------------------------------------------------------------------------------
def test_interpolate_kats(kats):  # interpolation.rs:681-720, 834-1008
    FIR5 = [0.019827, 0.132513, 0.347660, 0.347660, 0.132513, 0.019827]
    FIR12 = [-2.6551e-03, 1.5106e-04, 1.6104e-02, 5.9695e-02, 1.2705e-01, 1.9096e-01, 2.1739e-01, 1.9096e-01,
             1.2705e-01, 5.9695e-02, 1.6104e-02, 1.5106e-04, -2.6551e-03]
    d6 = np.zeros(6, dtype=np.complex64); d6[3] = 1
    d7 = np.zeros(7, dtype=np.complex64); d7[3] = 1
    got = DspVec(d6).interpolate(bd.SINC, 0.0, 12, 0.0).to_real().to_numpy()
    assert np.max(np.abs(got - vals(kats, "interpolate_sinc_even_test"))) < 1e-4
    got = DspVec(d7).interpolate(bd.SINC, 0.0, 14, 0.0).to_real().to_numpy()
    assert np.max(np.abs(got - vals(kats, "interpolate_sinc_odd_test"))) < 1e-4
    got = DspVec(d6).interpolate(bd.SINC, 0.0, 13, 0.0).to_real().to_numpy()
    assert np.max(np.abs(got - vals(kats, "interpolate_by_fractional_sinc_test"))) < 0.1
    r6 = np.zeros(6, dtype=np.float32); r6[3] = 1
    v = DspVec(r6).interpolate(bd.SINC, 0.0, 13, 0.0)
    assert not v.is_complex() and v.len() == 13
    assert np.max(np.abs(v.to_numpy() - vals(kats, "interpolate_by_fractional_sinc_real_data_test"))) < 0.1
    got = DspVec(np.array(FIR5, dtype=np.complex64)).interpolate(bd.SINC, 0.0, 12, 1.0).magnitude().to_numpy()
    assert np.max(np.abs(got - vals(kats, "interpolate_delayed_sinc_test"))) < 0.1
    got = DspVec(np.array(FIR5, dtype=np.float32)).interpft(6).to_numpy()
    assert np.max(np.abs(got - vals(kats, "interpolate_identity"))) < 0.1
    got = DspVec(np.array(FIR12, dtype=np.complex64)).interpolate(bd.SINC, 0.0, 6, 0.0).magnitude().to_numpy()
    assert np.max(np.abs(got - vals(kats, "decimate_with_interpolate_test"))) < 1e-4


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_interpolate_vs_oracle(dtype):
    rng = np.random.default_rng(77)
    cases = [(1000, 2500, True, 0.0, "rc"), (777, 1000, True, 1.5, "rc"), (512, 2048, False, 0.0, "sinc"),
             (1001, 333, True, 0.25, "sinc"), (4096, 1024, False, 0.0, None), (300, 300, True, 2.0, None),
             (5000, 65536, True, 0.0, None), (1 << 16, 1 << 18, False, 3.0, "rc")]
    for n, dest, cplx_, delay, kind in cases:
        x = rand_c(rng, n, dtype) if cplx_ else rng.uniform(-10, 10, n).astype(dtype)
        delta = 0.5
        if kind == "rc":
            f = lambda t: o.raised_cosine_freq(t, 0.35, dtype)
            got = DspVec(x, delta=delta).interpolate(bd.RAISED_COSINE, 0.35, dest, delay)
        elif kind == "sinc":
            f = lambda t: o.sinc_freq(t, dtype)
            got = DspVec(x, delta=delta).interpolate(bd.SINC, 0.0, dest, delay)
        else:
            f = None
            got = DspVec(x, delta=delta).interpolate(None, 0.0, dest, delay) if delay == 0 else \
                DspVec(x, delta=delta).interpolate(lambda t: 1.0, 0.0, dest, delay)
            if delay != 0:
                f = lambda t: dtype(1.0)
        ref = o.interpolate(x, f, dest, delay, dtype, delta=delta)
        assert got.is_complex() == cplx_ and got.points() == dest
        assert o.rel_l2(got.to_numpy(), ref) <= 4 * tol(max(n, dest), dtype), (n, dest, kind)
        assert abs(got.delta() - delta / (dtype(dest) / dtype(n))) <= 1e-6 * delta
    # band-limited signal: interpft reproduces the signal on the finer grid
    n, m = 4096, 10240
    t = np.arange(n) / n
    x = (np.cos(2 * np.pi * 30 * t) + 0.5 * np.sin(2 * np.pi * 71 * t)).astype(dtype)
    tm = np.arange(m) / m
    want = np.cos(2 * np.pi * 30 * tm) + 0.5 * np.sin(2 * np.pi * 71 * tm)
    assert np.max(np.abs(DspVec(x).interpft(m).to_numpy() - want)) < (1e-4 if dtype == np.float32 else 1e-11)
    # real vector + asymmetric callback is rejected (ArgumentFunctionMustBeSymmetric = 10)
    v = DspVec(x)
    cb = getattr(bd.lib(), "RealFn" + v._s)(lambda _d, t: 1.0)
    assert v.result_code_of("interpolate_custom", cb, None, 0, m, 0.0) == 10


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_symmetric_transforms(dtype):  # tests/real_test.rs:581-605
    rng = np.random.default_rng(201511210)
    for n in (1001, 7, 4097, 65537):
        x = rng.uniform(-10, 10, n).astype(dtype)
        S = DspVec(x).plain_sfft()
        assert S.is_complex() and S.points() == (n + 1) // 2 and S.domain() == bd.FREQ
        full = DspVec(x).to_complex().plain_fft().to_numpy()
        mirrored = DspVec(x).plain_sfft().mirror()
        assert mirrored.points() == n
        assert o.rel_l2(mirrored.to_numpy(), full) <= tol(n, dtype)              # "Different FFT paths must equal"
        assert o.rel_l2(S.to_numpy(), o.plain_sfft(x)) <= tol(n, dtype)
        back = DspVec(x).plain_sfft().plain_sifft()
        assert not back.is_complex() and back.len() == n and back.domain() == bd.TIME
        assert o.rel_l2(back.to_numpy() / n, x) <= 2 * tol(n, dtype)             # "Ifft must give back the original"
        assert o.rel_l2(DspVec(x).sfft().to_numpy(), o.sfft(x)) <= tol(n, dtype)
        assert o.rel_l2(DspVec(x).windowed_sfft(bd.HAMMING).to_numpy(), o.windowed_sfft(x, bd.HAMMING, dtype)) <= 2 * tol(n, dtype)
    # sifft / windowed_sifft on a half spectrum whose shifted first bin is real
    p = 33
    H = rand_c(rng, p, dtype)
    Hs = H.copy(); Hs[(p // 2)] = Hs[p // 2].real      # ifft_shift brings element p//2 to the front
    ref = o.sifft(Hs)
    got = DspVec(Hs, domain=bd.FREQ).sifft()
    assert not isinstance(ref, int) and got.len() == 2 * p - 1
    assert o.rel_l2(got.to_numpy(), ref) <= 4 * tol(2 * p, dtype)
    got = DspVec(Hs, domain=bd.FREQ).windowed_sifft(bd.HAMMING).to_numpy()
    assert o.rel_l2(got, o.windowed_sifft(Hs, bd.HAMMING, dtype)) <= 4 * tol(2 * p, dtype)
    # error behaviour
    assert DspVec(np.ones(8, dtype=dtype)).result_code_of("plain_sfft") == 9          # even length
    assert DspVec(np.ones(7, dtype=dtype).astype(np.complex64 if dtype == np.float32 else np.complex128)).result_code_of("sfft") == 5
    bad = rand_c(rng, 9, dtype); bad[0] = 1 + 1j
    assert DspVec(bad, domain=bd.FREQ).result_code_of("plain_sifft") == 8           # not conj-symmetric
    assert DspVec(bad, domain=bd.TIME).result_code_of("plain_sifft") == 6


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_multiply_complex_exponential(dtype):  # complex_ops.rs:81-105
    rng = np.random.default_rng(3)
    x = rand_c(rng, 5000, dtype)
    got = DspVec(x, delta=0.5).multiply_complex_exponential(0.01, 0.3).to_numpy()
    assert o.rel_l2(got, o.multiply_complex_exponential(x, 0.01, 0.3, dtype, delta=0.5)) <= 1e-6 if dtype == np.float32 else 1e-14


@pytest.mark.parametrize("log2n", [21, 22, 23, 24])
def test_three_pass_transforms(log2n):
    """packed three-pass path (fftp.cu): forward (plain, shifted + magnitude) and inverse of one long sequence."""
    n = 1 << log2n
    rng = np.random.default_rng(log2n)
    L = bd.lib()
    rows = 2 if log2n < 23 else 1
    x = rand_c(rng, n * rows, np.float32)
    v = DspVec(x)
    out = DspVec.zeros(2 * n * rows, is_complex=True, dtype=np.float32)
    pin, pout = v._fn("bdsp_device_ptr")(v._h), out._fn("bdsp_device_ptr")(out._h)
    xr = x.reshape(rows, n).astype(np.complex128)
    X = np.fft.fft(xr, axis=1)
    assert L.bdsp_fft_rows_c32(pin, pout, n, rows, 0) == 0
    assert o.rel_l2(out.to_numpy().reshape(rows, n), X) <= tol(n, np.float32)
    assert L.bdsp_fft_rows_c32(pin, pout, n, rows, bd.F_INVERSE) == 0
    assert o.rel_l2(out.to_numpy().reshape(rows, n), np.fft.ifft(xr, axis=1) * n) <= tol(n, np.float32)
    assert L.bdsp_fft_rows_c32(pin, pout, n, rows, bd.F_INVERSE | bd.F_SHIFT) == 0
    assert o.rel_l2(out.to_numpy().reshape(rows, n), np.fft.ifft(np.fft.ifftshift(xr, axes=1), axis=1)) <= tol(n, np.float32)
    mag = DspVec.zeros(n * rows, dtype=np.float32)
    assert L.bdsp_fft_rows_c32(pin, mag._fn("bdsp_device_ptr")(mag._h), n, rows, bd.F_SHIFT | bd.F_MAGNITUDE) == 0
    assert o.rel_l2(mag.to_numpy().reshape(rows, n), np.abs(np.fft.fftshift(X, axes=1))) <= tol(n, np.float32)


def test_streams_do_not_share_workspace():
    """Calls queued on different streams may overlap on the device: their scratch buffers must be distinct."""
    rng = np.random.default_rng(99)
    L = bd.lib()
    n = 1 << 18
    xs = [rand_c(rng, n, np.float32) for _ in range(3)]
    streams = [L.bdsp_stream_create() for _ in xs]
    vecs = [DspVec(x) for x in xs]
    for _ in range(3):                       # several rounds in flight on every stream before anything is synchronised
        for st, v in zip(streams, vecs):
            L.bdsp_set_stream(st)
            v.plain_fft()
            v.plain_ifft()
            v.scale(1.0 / n)
    for st in streams:
        L.bdsp_stream_sync(st)
    L.bdsp_set_stream(None)
    for v, x in zip(vecs, xs):
        assert o.rel_l2(v.to_numpy(), x) <= 8 * tol(n, np.float32)
    for st in streams:
        L.bdsp_stream_destroy(st)


@pytest.mark.parametrize("log2n", [20, 24])
def test_large_transform_properties(log2n):
    """size-independent properties at the BASELINE / maximum sizes: Parseval, linearity, fft -> ifft round trip."""
    n = 1 << log2n
    rng = np.random.default_rng(1000 + log2n)
    x = rand_c(rng, n, np.float32)
    y = rand_c(rng, n, np.float32)
    X = DspVec(x).plain_fft().to_numpy().astype(np.complex128)
    Y = DspVec(y).plain_fft().to_numpy().astype(np.complex128)
    e_time = float(np.sum(np.abs(x.astype(np.complex128)) ** 2))
    assert abs(float(np.sum(np.abs(X) ** 2)) / n - e_time) <= 1e-5 * e_time                     # Parseval
    Z = DspVec((x + 2 * y).astype(np.complex64)).plain_fft().to_numpy().astype(np.complex128)
    assert o.rel_l2(Z, X + 2 * Y) <= 2 * tol(n, np.float32)                                    # linearity
    back = DspVec(x).fft().ifft().to_numpy()
    assert o.rel_l2(back, x) <= 2 * tol(n, np.float32)                                         # round trip
    # a pure tone lands in one bin
    k0 = 12345
    tone = np.exp(2j * np.pi * k0 * np.arange(n) / n).astype(np.complex64)
    T = np.abs(DspVec(tone).plain_fft().to_numpy())
    assert int(np.argmax(T)) == k0 and abs(T[k0] / n - 1) < 1e-3
