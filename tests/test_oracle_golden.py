"""Pins oracle/dsp_oracle.py against the known-answer vectors of the reference's own tests.

Expected values come from tests/golden/reference_kats.json (extracted from the reference sources by
tests/golden/extract_goldens.py); the input construction of every case is restated here with the
reference file:line it follows.  CPU only."""
import math
import numpy as np
import pytest

from oracle import dsp_oracle as o


def vals(kats, key):
    return np.array(kats[key]["values"], dtype=np.float64)


def cplx(interleaved):
    a = np.asarray(interleaved, dtype=np.float64)
    return a[0::2] + 1j * a[1::2]


# --- transforms ------------------------------------------------------------------------------
def test_doc_plain_fft(kats):  # time_to_freq.rs:26-39
    k = kats["doc_plain_fft"]
    assert np.allclose(o.plain_fft(cplx(k["input"])), cplx(k["values"]), atol=1e-4)


def test_doc_fft(kats):  # time_to_freq.rs:48-61 (odd length shift)
    k = kats["doc_fft"]
    assert np.allclose(o.fft(cplx(k["input"])), cplx(k["values"]), atol=1e-4)


def test_doc_plain_ifft(kats):  # freq_to_time.rs:28-41
    k = kats["doc_plain_ifft"]
    assert np.allclose(o.plain_ifft(cplx(k["input"])), cplx(k["values"]), atol=1e-4)


def test_doc_ifft(kats):  # freq_to_time.rs:50-63
    k = kats["doc_ifft"]
    assert np.allclose(o.ifft(cplx(k["input"])), cplx(k["values"]), atol=1e-4)


def sinusoid64():  # tests/time_freq_test.rs:221-231
    n = np.arange(64, dtype=np.float64)
    return np.cos(n * 0.1 * 2.0 * np.pi + 0.25)


def test_fft_vector64(kats):  # tests/time_freq_test.rs:45-120, Octave-generated
    got = o.magnitude(o.fft(sinusoid64()), np.float64)
    assert np.max(np.abs(got - vals(kats, "fft_vector64"))) < 1e-6


def test_fft_ifft_vector64():  # tests/time_freq_test.rs:199-207
    x = sinusoid64()
    assert np.allclose(o.ifft(o.fft(x)).real, x, atol=1e-12)


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 8, 9, 10, 11])
def test_swap_halves_literal_vs_numpy(n):  # vector_types/mod.rs:171-191, tests :693-712
    data = list(range(1, n + 1))
    assert o.swap_array_halves_literal(data, True) == list(np.fft.fftshift(data))
    if n > 1:  # the reference indexes out of bounds for the inverse swap of a single point
        assert o.swap_array_halves_literal(data, False) == list(np.fft.ifftshift(data))
    assert list(o.fft_shift(np.array(data))) == list(np.fft.fftshift(data))
    assert list(o.ifft_shift(np.array(data))) == list(np.fft.ifftshift(data))


def test_swap_halves_reference_cases():  # vector_types/mod.rs:693-712
    assert o.swap_array_halves_literal([1, 2, 3, 4], True) == [3, 4, 1, 2]
    assert o.swap_array_halves_literal([1, 2, 3, 4, 5], True) == [4, 5, 1, 2, 3]
    assert o.swap_array_halves_literal([1, 2, 3, 4, 5], False) == [3, 4, 5, 1, 2]


# --- impulse / frequency responses ---------------------------------------------------------------
def conv_test(f, n, step):  # conv_types.rs:522-541
    j0 = -(n // 2)
    return np.array([f((j0 + i) * step) for i in range(n)], dtype=np.float64)


def test_raised_cosine(kats):  # conv_types.rs:582-598
    e = vals(kats, "raised_cosine_test")
    got = conv_test(lambda x: o.raised_cosine_impulse(x, 0.35, np.float64), len(e), 0.2)
    assert np.max(np.abs(got - e)) < 1e-4


def test_sinc(kats):  # conv_types.rs:600-607
    e = vals(kats, "sinc_test")
    got = conv_test(lambda x: o.sinc_impulse(np.float32(x), np.float32), len(e), np.float32(0.5))
    assert np.max(np.abs(got - e)) < 1e-4


def test_sinc_freq(kats):  # conv_types.rs:609-614
    e = vals(kats, "sinc_freq_test")
    got = conv_test(lambda x: o.sinc_freq(x, np.float32), len(e), 0.5)
    assert np.max(np.abs(got - e)) < 1e-4


def test_rc_freq(kats):  # conv_types.rs:686-702
    e = vals(kats, "freq_test")
    got = conv_test(lambda x: o.raised_cosine_freq(x, 0.5, np.float64), len(e), 0.4)
    assert np.max(np.abs(got - e)) < 0.1


def test_rc_special_point():  # conv_types.rs:411-414: |x| == 1/(2 beta)
    v = o.raised_cosine_impulse(1.0, 0.5, np.float64)
    assert abs(v - np.sin(np.pi) / np.pi * np.pi / 4) < 1e-12
    assert np.isfinite(v)


def test_fft_swap_x(kats):  # time_freq/mod.rs:854-863
    inp = [-4.0, -3.0, -2.0, -1.0, 0.0, 1.0, 2.0, 3.0, 4.0]
    got = [o.fft_swap_x(True, x, 4.0) for x in inp]
    assert got == kats["fft_swap_x_test"]["values"]


# --- convolution -------------------------------------------------------------------------------------
def test_multiply_frequency_response_odd(kats):  # convolution.rs:632-639 (5 points)
    X = np.ones(5, dtype=np.complex128) * (1 + 1j)
    got = o.multiply_frequency_response(X, lambda x: o.raised_cosine_freq(x, 1.0, np.float32), 2.0, np.float32)
    flat = np.empty(10)
    flat[0::2], flat[1::2] = got.real, got.imag
    assert np.max(np.abs(flat - vals(kats, "convolve_complex_freq_and_freq32"))) < 1e-4


def test_multiply_frequency_response_even(kats):  # convolution.rs:641-648 (6 points)
    X = np.ones(6, dtype=np.complex128) * (1 + 1j)
    got = o.multiply_frequency_response(X, lambda x: o.raised_cosine_freq(x, 1.0, np.float32), 2.0, np.float32)
    flat = np.empty(12)
    flat[0::2], flat[1::2] = got.real, got.imag
    assert np.max(np.abs(flat - vals(kats, "convolve_complex_freq_and_freq_even32"))) < 1e-4


def test_convolve_real_rc(kats):  # convolution.rs:650-669
    x = np.zeros(10)
    x[5] = 1.0
    got = o.convolve_function(x, lambda t: o.raised_cosine_impulse(t, 0.35, np.float32), 0.2, 5, np.float32)
    assert np.max(np.abs(got - vals(kats, "convolve_real_time_and_time32"))) < 1e-4


def test_convolve_complex_sinc(kats):  # convolution.rs:671-702
    n = 11
    x = np.zeros(n, dtype=np.complex128)
    x[n // 2] = 1.0  # data_mut(len) = 1.0 -> real part of point 5
    got = o.convolve_function(x, lambda t: o.sinc_impulse(t, np.float32), 0.5, n // 2, np.float32)
    assert np.max(np.abs(np.abs(got) - vals(kats, "convolve_complex_time_and_time32"))) < 1e-4


def test_convolve_complex_vectors(kats):  # convolution.rs:737-775
    n = 11
    x = np.zeros(n, dtype=np.complex128)
    x[n // 2] = 1.0
    h = np.array([o.sinc_impulse(np.float32(v * 0.5), np.float32) for v in np.arange(-5.0, 6.0)], dtype=np.float64)
    got = o.convolve_signal_direct(x, h.astype(np.complex128))
    assert np.max(np.abs(np.abs(got) - vals(kats, "convolve_complex_vectors32"))) < 1e-4
    assert np.allclose(got, o.convolve_signal(x, h.astype(np.complex128)), atol=1e-12)


def test_shift_as_conv(kats):  # convolution.rs:818-842
    a = np.arange(10, dtype=np.float64).astype(np.complex128)
    b = np.zeros(10, dtype=np.complex128)
    b[4] = 1.0
    assert np.allclose(np.abs(o.convolve_signal_direct(a, b)), vals(kats, "shift_left_by_1_as_conv"), atol=1e-4)
    b = np.array([0.0, 0.0, 1.0], dtype=np.complex128)
    assert np.allclose(np.abs(o.convolve_signal_direct(a, b)), vals(kats, "shift_left_by_1_as_conv_shorter"), atol=1e-4)


@pytest.mark.parametrize("n", [10, 9])
def test_conv_vs_freq_multiplication(n):  # convolution.rs:802-882
    a = np.arange(n, dtype=np.float64).astype(np.complex128)
    b = (15.0 - np.arange(n, dtype=np.float64)).astype(np.complex128)
    conv = o.convolve_signal_direct(a, b)
    m = o.ifft(o.mul(o.fft(a), o.fft(b), np.float64))
    m = o.swap_halves(m[::-1])
    assert np.allclose(np.abs(m), np.abs(conv), atol=1e-4)
    if n % 2 == 0:
        assert np.allclose(m, conv, atol=1e-4)


def test_overlap_discard_case():  # convolution.rs:884-898: direct == FFT identity
    a = np.arange(100, dtype=np.float64).astype(np.complex128)
    b = np.array([0.1, 0.2, 0.3, 0.5, 0.1, 0.2], dtype=np.complex128)
    assert np.allclose(o.convolve_signal_direct(a, b), o.convolve_signal(a, b), atol=1e-10)


def test_convolve_signal_random_direct_vs_fft():
    rng = np.random.default_rng(7)
    for n, l in [(64, 1), (64, 2), (65, 7), (257, 64), (300, 300), (1000, 129)]:
        x = rng.uniform(-10, 10, n) + 1j * rng.uniform(-10, 10, n)
        h = rng.uniform(-1, 1, l) + 1j * rng.uniform(-1, 1, l)
        assert o.rel_l2(o.convolve_signal(x, h), o.convolve_signal_direct(x, h)) < 1e-13


def test_convolve_signal_error_codes():  # convolution.rs:485-492
    assert o.check_convolve_signal_args(10, True, 0, 1.0, 11, True, 0, 1.0) == o.ERR_INVALID_ARG_LEN
    assert o.check_convolve_signal_args(10, True, 1, 1.0, 5, True, 1, 1.0) == o.ERR_MUST_BE_TIME
    assert o.check_convolve_signal_args(10, True, 0, 1.0, 5, False, 0, 1.0) == o.ERR_META_DATA
    assert o.check_convolve_signal_args(10, True, 0, 1.0, 5, True, 0, 2.0) == o.ERR_META_DATA
    assert o.check_convolve_signal_args(10, True, 0, 1.0, 5, True, 0, 1.05) == o.ERR_OK


def test_convolve_simd_branch_equals_function_branch_for_symmetric():
    # tests/convolution_test.rs:73-111 (optimised vs not optimised), ratio = 1
    rng = np.random.default_rng(3)
    x = rng.uniform(-10, 10, 1500) + 1j * rng.uniform(-10, 10, 1500)
    f = lambda t: o.raised_cosine_impulse(t, 0.35, np.float32)
    assert o.would_benefit_from_simd(3000, 12, 1.0, np.float32)
    a = o.convolve_function(x, f, 1.0, 12, np.float32)              # SIMD branch
    b = o.convolve_function(x, f, 1.0, 12, np.float32, len_in_T=100)  # forced function branch
    assert o.rel_l2(a, b) < 1e-12


# --- interpolation ---------------------------------------------------------------------------------------
def _sinc32(t):
    return o.sinc_impulse(t, np.float32)


def test_interpolatef_integer_even(kats):  # interpolation.rs:752-773
    n = 6
    x = np.zeros(n, dtype=np.complex128)
    x[n // 2] = 1.0
    got = o.interpolatef(x, _sinc32, 2.0, 0.0, n, np.float32).real
    assert np.max(np.abs(got - vals(kats, "interpolatef_by_integer_sinc_even_test"))) < 0.1


def test_interpolatef_integer_odd(kats):  # interpolation.rs:775-796
    n = 7
    x = np.zeros(n, dtype=np.complex128)
    x[n // 2] = 1.0
    got = o.interpolatef(x, _sinc32, 2.0, 0.0, n, np.float32).real
    assert np.max(np.abs(got - vals(kats, "interpolatef_by_integer_sinc_odd_test"))) < 0.1


def test_interpolatef_fractional(kats):  # interpolation.rs:798-831
    n = 6
    x = np.zeros(n, dtype=np.complex128)
    x[n // 2] = 1.0
    got = o.interpolatef(x, _sinc32, np.float32(13.0 / 6.0), 0.0, n, np.float32).real
    e = vals(kats, "interpolatef_by_fractional_sinc_test")
    assert len(got) == len(e)
    assert np.max(np.abs(got - e)) < 0.1


def test_interpolatef_delayed(kats):  # interpolation.rs:899-919
    n = 6
    x = np.zeros(n, dtype=np.complex128)
    x[n // 2] = 1.0
    got = np.abs(o.interpolatef(x, _sinc32, 2.0, 1.0, n, np.float32))
    assert np.max(np.abs(got - vals(kats, "interpolatef_delayed_sinc_test"))) < 0.1


def test_interpolate_lin(kats):  # real_interpolation.rs:226-237
    x = np.array([-1.0, -2.0, -1.0, 0.0, 1.0, 3.0, 4.0])
    got = o.interpolate_lin(x, 4.0, 0.0, np.float64)
    e = vals(kats, "linear_test")
    assert len(got) == len(e)
    assert np.max(np.abs(got - e)) < 1e-12


def test_interpolate_lin_counter_saturation_flag():
    x = np.arange(8, dtype=np.float32)
    a = o.interpolate_lin(x, 2.0, 0.0, np.float32, replicate_counter_saturation=False)
    b = o.interpolate_lin(x, 2.0, 0.0, np.float32, replicate_counter_saturation=True)
    assert np.array_equal(a, b)  # identical below 2^24


# --- literal transliteration of the SIMD interpolation path vs the closed form -------------------
def _create_shifted_copies_literal(vec, is_complex, reg_len):
    """time_freq/mod.rs:81-165 restated literally on Python lists (T scalars, interleaved)."""
    step = 2 if is_complex else 1
    number_of_shifts = reg_len // step
    copies = []
    n = len(vec)
    for i in range(number_of_shifts):
        rev = list(vec[::-1])
        it = iter(rev)
        shift = ((number_of_shifts - i) % number_of_shifts) * step
        min_len = n + shift
        ln = (min_len + reg_len - 1) // reg_len
        copy = []
        j = ln * reg_len
        cur = []
        while j > 0:
            j -= step
            if j < shift or j >= min_len:
                cur.extend([0.0] * step)
            elif step > 1:
                im = next(it)
                re = next(it)
                cur.extend([re, im])
            else:
                cur.append(next(it))
            if len(cur) >= reg_len:
                copy.append(cur[:reg_len])
                cur = cur[reg_len:]
        copies.append(copy)
    return copies


def _interp_simd_interior_literal(x, vs, F, L, reg_len):
    """interpolation.rs:203-275 restated for REAL data (step = 1, left_points = 0)."""
    shifts = []
    for s in range(F):
        shifts.extend(_create_shifted_copies_literal(list(vs[s]), False, reg_len))
    n = len(x)
    pad = (-n) % reg_len
    data = list(x) + [0.0] * pad
    regs = [data[k:k + reg_len] for k in range(0, len(data), reg_len)]
    new_points = n * F
    scalar_len = (2 * L + 1) * F
    out = {}
    for i in range(scalar_len, new_points - scalar_len):
        rounded = (i + F - 1) // F
        end = rounded + L
        simd_end = (end + reg_len - 1) // reg_len
        simd_shift = end % reg_len
        factor_shift = (F - i % F) % F
        shifted = shifts[factor_shift * reg_len + simd_shift]
        acc = 0.0
        for reg, other in zip(regs[simd_end - len(shifted):simd_end], shifted):
            acc += sum(a * b for a, b in zip(reg, other))
        out[i] = acc
    return out


@pytest.mark.parametrize("reg_len", [4, 2])
@pytest.mark.parametrize("delay", [0.0, 0.3])
def test_interpolatef_closed_form_matches_literal_simd_path(reg_len, delay):
    rng = np.random.default_rng(11)
    n, F, L = 700, 3, 5
    x = rng.uniform(-10, 10, n)
    f = lambda t: o.raised_cosine_impulse(t, 0.35, np.float64)
    got = o.interpolatef(x, f, float(F), delay, L, np.float64)
    assert len(got) == n * F
    vs = o.interpolatef_tap_vectors(f, L, F, delay, np.float64)
    lit = _interp_simd_interior_literal(x, vs, F, L, reg_len)
    idx = np.array(sorted(lit))
    ref = np.array([lit[i] for i in idx])
    assert np.max(np.abs(got[idx] - ref)) < 1e-10


def test_interpolatef_fast_vs_scalar_path_delay0():
    # tests/interpolation_test.rs pattern: SIMD path vs scalar path agree for symmetric f, delay 0,
    # except for the differing truncation windows (tail taps) -> loose tolerance like the reference.
    rng = np.random.default_rng(5)
    n, F, L = 800, 4, 12
    x = np.convolve(rng.uniform(-1, 1, n), np.ones(8) / 8, mode="same")  # band-limited-ish
    fast = o.interpolatef(x, _sinc32, 4.0, 0.0, L, np.float32)
    assert o.interpolatef_uses_fast_path(L, n * F, 4.0, np.float32)
    # decimating the interpolated signal returns the input (sinc is 0 at non-zero integers)
    assert np.max(np.abs(fast[::F] - x)) < 1e-5


# --- elementwise ------------------------------------------------------------------------------------
def test_magnitude_phase_doc():  # complex_to_real.rs:24-34,88-99
    x = np.array([3.0 + -4.0j, -3.0 + 4.0j])
    assert np.array_equal(o.magnitude(x, np.float64), [5.0, 5.0])
    x = np.array([1.0 + 0j, 0 + 4.0j, -2.0 + 0j, 0 - 3.0j, 1.0 - 1.0j])
    e = [0.0, np.pi / 2, np.pi, -np.pi / 2, -np.pi / 4]
    assert np.allclose(o.phase(x, np.float64), e, atol=1e-12)


def test_mul_no_fma_semantics():
    rng = np.random.default_rng(1)
    a = (rng.uniform(-10, 10, 1000) + 1j * rng.uniform(-10, 10, 1000)).astype(np.complex64)
    b = (rng.uniform(-10, 10, 1000) + 1j * rng.uniform(-10, 10, 1000)).astype(np.complex64)
    got = o.mul(a, b, np.float32)
    exact = a.astype(np.complex128) * b.astype(np.complex128)
    assert o.rel_l2(got, exact) < 1e-6
    assert o.ulp_diff(np.float32([1.0]), np.nextafter(np.float32(1.0), np.float32(2.0)), np.float32)[0] == 1
    assert o.ulp_diff(np.float64([-1.0]), np.float64([-1.0]), np.float64)[0] == 0


# --- windows and correlation -------------------------------------------------------
@pytest.mark.parametrize("kind,key", [(0, "triangular_window32_test"), (1, "hamming_window32_test"), (2, "blackmanharris_window32_test")])
def test_window_functions(kats, kind, key):  # window_functions.rs:156-175
    e = vals(kats, key)
    got = np.array([o.window_value(kind, i, len(e), np.float32) for i in range(len(e))], dtype=np.float64)
    assert np.max(np.abs(got - e)) < 1e-4
    assert np.allclose(o.window_table(kind, len(e), np.float32), got, atol=1e-7)
    assert np.all(o.window_table(3, 7, np.float32) == 1.0)


def test_windowed_fft_vector64(kats):  # tests/time_freq_test.rs:122-197
    got = o.magnitude(o.windowed_fft(sinusoid64(), 1, np.float64), np.float64)
    assert np.max(np.abs(got - vals(kats, "windowed_fft_vector64"))) < 1e-6


def test_windowed_fft_ifft_roundtrip():  # tests/time_freq_test.rs:209-219
    x = sinusoid64()
    assert np.allclose(o.windowed_ifft(o.windowed_fft(x, 1, np.float64), 1, np.float64).real, x, atol=1e-10)


def test_time_correlation(kats):  # correlation.rs:170-196
    a = cplx(kats["time_correlation_test_a"]["values"])
    b = cplx(kats["time_correlation_test_b"]["values"])
    c = vals(kats, "time_correlation_test_c")
    got = o.correlate(a, o.prepare_argument_padded(b))
    flat = np.empty(2 * len(got)); flat[0::2], flat[1::2] = got.real, got.imag
    assert len(flat) == len(c) and np.max(np.abs(flat - c)) < 0.1


def test_time_correlation2(kats):  # correlation.rs:198-215
    a = np.array([1 + 1j, 2 + 1j, 3 + 1j])
    b = np.array([4 + 1j, 5 + 1j, 6 + 1j])
    c = vals(kats, "time_correlation_test2_c")
    got = o.correlate(a, o.prepare_argument_padded(b))
    flat = np.empty(2 * len(got)); flat[0::2], flat[1::2] = got.real, got.imag
    assert np.max(np.abs(flat - c)) < 0.1
    # definition check: full cross-correlation sum_n a[n + k] conj(b[n])
    ref = np.correlate(a, b, mode="full")
    assert np.allclose(got, ref, atol=1e-10)


def test_interpolatei_kats(kats):  # interpolation.rs:653-680 (sinc), :722-750 (raised cosine 0.4)
    n = 6
    x = np.zeros(n, dtype=np.complex128)
    x[n // 2] = 1.0
    got = np.abs(o.interpolatei(x, lambda t: o.sinc_freq(t, np.float32), 2, np.float32))
    assert np.max(np.abs(got - vals(kats, "interpolatei_sinc_test"))) < 1e-4
    got = np.abs(o.interpolatei(x, lambda t: o.raised_cosine_freq(t, 0.4, np.float32), 2, np.float32))
    assert np.max(np.abs(got - vals(kats, "interpolatei_rc_test"))) < 1e-4


# --------------------------------------------------------------------------------------------------
# FFT-based resampling (interpolate / interpft) and the symmetric transforms
# --------------------------------------------------------------------------------------------------
FIR5 = [0.019827, 0.132513, 0.347660, 0.347660, 0.132513, 0.019827]              # interpolation.rs:914
FIR12 = [-2.6551e-03, 1.5106e-04, 1.6104e-02, 5.9695e-02, 1.2705e-01, 1.9096e-01, 2.1739e-01, 1.9096e-01,
         1.2705e-01, 5.9695e-02, 1.6104e-02, 1.5106e-04, -2.6551e-03]           # interpolation.rs:975-989


def _dirac(n, cplx=True):
    x = np.zeros(n, dtype=np.complex128 if cplx else np.float64)
    x[n // 2] = 1.0
    return x


def test_interpolate_kats(kats):  # interpolation.rs:681-720, 834-910, 912-960, 962-1008
    f32 = np.float32
    sinc = lambda t: o.sinc_freq(t, f32)
    got = o.interpolate(_dirac(6), sinc, 12, 0.0, f32).real
    assert np.max(np.abs(got - vals(kats, "interpolate_sinc_even_test"))) < 1e-4
    got = o.interpolate(_dirac(7), sinc, 14, 0.0, f32).real
    assert np.max(np.abs(got - vals(kats, "interpolate_sinc_odd_test"))) < 1e-4
    got = o.interpolate(_dirac(6), sinc, 13, 0.0, f32).real
    assert np.max(np.abs(got - vals(kats, "interpolate_by_fractional_sinc_test"))) < 0.1
    got = o.interpolate(_dirac(6, cplx=False), sinc, 13, 0.0, f32)
    assert not np.iscomplexobj(got)
    assert np.max(np.abs(got - vals(kats, "interpolate_by_fractional_sinc_real_data_test"))) < 0.1
    got = np.abs(o.interpolate(np.array(FIR5, dtype=np.complex128), sinc, 12, 1.0, f32))
    assert np.max(np.abs(got - vals(kats, "interpolate_delayed_sinc_test"))) < 0.1
    got = o.interpft(np.array(FIR5), 6, f32)
    assert np.max(np.abs(got - vals(kats, "interpolate_identity"))) < 0.1
    got = np.abs(o.interpolate(np.array(FIR12, dtype=np.complex128), sinc, 6, 0.0, f32))
    assert np.max(np.abs(got - vals(kats, "decimate_with_interpolate_test"))) < 1e-4


def test_interpft_is_band_limited_resampling():
    """interpft of a band-limited periodic signal reproduces the signal on the finer grid (the property behind
    the reference's Octave interpft comparisons, interpolation.rs:920-960)."""
    n, m = 64, 160
    t = np.arange(n) / n
    x = np.cos(2 * np.pi * 3 * t) + 0.5 * np.sin(2 * np.pi * 7 * t)
    tm = np.arange(m) / m
    want = np.cos(2 * np.pi * 3 * tm) + 0.5 * np.sin(2 * np.pi * 7 * tm)
    assert np.max(np.abs(o.interpft(x, m, np.float64) - want)) < 1e-12
    # and an integer delay is a circular advance
    got = o.interpolate(x.astype(np.complex128), None, n, 2.0, np.float64)
    assert np.max(np.abs(got.real - np.roll(x, -2))) < 1e-12


def test_symmetric_transforms():  # tests/real_test.rs:581-605 (real_fft_test32)
    rng = np.random.default_rng(201511210)
    x = rng.uniform(-10, 10, 1001)
    S = o.plain_sfft(x)
    assert len(S) == 501
    assert o.rel_l2(o.mirror(S), np.fft.fft(x)) < 1e-12            # "Different FFT paths must equal"
    back = o.plain_sifft(S) / 1001.0
    assert np.max(np.abs(back - x)) < 1e-9                         # "Ifft must give back the original result"
    bad = S.copy(); bad[0] += 1j
    assert o.plain_sifft(bad) == o.ERR_CONJ_SYMMETRIC
    assert len(o.sfft(x)) == 501 and np.allclose(o.sfft(x)[-1], np.sum(x))   # shifted: DC is the last kept bin


# --------------------------------------------------------------------------------------------------
# rest of the facade: the reference's doc examples and unit tests pin the oracle
# --------------------------------------------------------------------------------------------------
def test_hermite_kats():  # real_interpolation.rs:196-228
    got = o.interpolate_hermite([-1.0, -2.0, -1.0, 0.0, 1.0, 3.0, 4.0], 4.0, 0.0, np.float32)
    exp = [-1.0000, -1.4375, -1.7500, -1.9375, -2.0000, -1.8906, -1.6250, -1.2969, -1.0000, -0.7500, -0.5000, -0.2500, 0.0,
           0.2344, 0.4583, 0.7031, 1.0000, 1.4375, 2.0000, 2.5625, 3.0000, 3.3203, 3.6042, 3.8359, 4.0]
    assert len(got) == len(exp) and np.max(np.abs(got[4:-4] - np.array(exp)[4:-4])) < 6e-2
    got = o.interpolate_hermite([-3.0, -2.0, -1.0, 0.0, 1.0, 2.0, 3.0], 3.0, 0.0, np.float32)
    assert np.max(np.abs(got - np.linspace(-3, 3, 19))) < 5e-3


def test_statistics_doc_examples():  # statistics.rs:45-66, 82-92, 98-131
    z = np.array([1 + 2j, 3 + 4j, 5 + 6j], dtype=np.complex64)
    st = o.statistics(z)
    assert st["sum"] == 9 + 12j and st["count"] == 3 and st["average"] == 3 + 4j
    assert abs(st["rms"] - (3.4027193 + 4.3102784j)) < 1e-4
    assert (st["min"], st["min_index"], st["max"], st["max_index"]) == (1 + 2j, 0, 5 + 6j, 2)
    sp = o.statistics_split(z, 2)
    assert sp[0]["sum"] == 6 + 8j and sp[1]["sum"] == 3 + 4j
    assert np.sum(z.astype(np.complex128) ** 2) == -21 + 88j
    assert o.statistics_split(z, 17) == o.ERR_INVALID_ARG_LEN


def test_elementwise_doc_examples():  # trigonometry_and_powers.rs:9-191, diff_sum.rs:13-60, real_ops.rs doc examples
    f64 = np.float64
    assert np.allclose(o.real_math("sin", [math.pi / 2, -math.pi / 2], f64), [1.0, -1.0])
    assert np.array_equal(o.real_math("sqrt", [1.0, 4.0, 9.0, 16.0, 25.0], f64), [1, 2, 3, 4, 5])
    assert np.array_equal(o.real_math("square", [1.0, 2.0, 3.0, 4.0, 5.0], f64), [1, 4, 9, 16, 25])
    assert np.allclose(o.real_math("root", [1.0, 8.0, 27.0], f64, 3.0), [1, 2, 3])
    assert np.allclose(o.real_math("powf", [1.0, 2.0, 3.0], f64, 3.0), [1, 8, 27])
    assert np.allclose(o.real_math("ln", [2.718281828459045, 7.389056, 20.085537], f64), [1, 2, 3], atol=1e-6)
    assert np.allclose(o.real_math("log", [10.0, 100.0, 1000.0], f64, 10.0), [1, 2, 3])
    assert np.allclose(o.real_math("expf", [1.0, 2.0, 3.0], f64, 10.0), [10, 100, 1000])
    assert np.array_equal(o.diff(np.array([2.0, 3.0, 2.0, 6.0])), [1.0, -1.0, 4.0])
    assert np.array_equal(o.diff(np.array([2.0, 3.0, 2.0, 6.0]), with_start=True), [2.0, 1.0, -1.0, 4.0])
    assert np.array_equal(o.cum_sum(np.array([2.0, 1.0, -1.0, 4.0])), [2.0, 3.0, 2.0, 6.0])
    assert np.array_equal(o.real_math("wrap", [1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0, 8.0], f64, 4.0), [1, 2, 3, 0, 1, 2, 3, 0])
    assert np.array_equal(o.unwrap(np.array([1.0, 2.0, 3.0, 0.0, 1.0, 2.0, 3.0, 0.0]), 4.0, f64), [1, 2, 3, 4, 5, 6, 7, 8])
    # complex functions against an independent implementation (NumPy's C99 complex math) in f64
    rng = np.random.default_rng(0)
    z = rng.uniform(-2, 2, 500) + 1j * rng.uniform(-2, 2, 500)
    for name, fn in [("sin", np.sin), ("cos", np.cos), ("sinh", np.sinh), ("cosh", np.cosh), ("asin", np.arcsin),
                     ("acos", np.arccos), ("atan", np.arctan), ("asinh", np.arcsinh), ("acosh", np.arccosh),
                     ("atanh", np.arctanh), ("sqrt", np.sqrt), ("ln", np.log), ("exp", np.exp)]:
        got = o.complex_math(name, z, f64)
        assert np.max(np.abs(got - fn(z)) / np.maximum(1, np.abs(fn(z)))) < 1e-13, name
    assert np.max(np.abs(o.complex_math("powf", z, f64, 2.5) - z ** 2.5)) < 1e-12
    assert np.max(np.abs(o.complex_math("expf", z, f64, 3.0) - 3.0 ** z)) < 1e-12


def test_split_merge_and_smaller():
    x = np.arange(12.0)
    parts = o.split_into(x, 3)
    assert [p.tolist() for p in parts] == [[0, 3, 6, 9], [1, 4, 7, 10], [2, 5, 8, 11]]
    assert np.array_equal(o.merge(parts), x)
    assert o.split_into(x, 5) == o.ERR_INVALID_ARG_LEN
    assert np.array_equal(o.binary_smaller("add", x, np.array([10.0, 20.0]), np.float64), x + np.tile([10.0, 20.0], 6))
