"""Small invocations of every kernel family (run under compute-sanitizer)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import basic_dsp_b200 as bd
from basic_dsp_b200 import DspVec

L = bd.lib(); bd.require_device()
rng = np.random.default_rng(0)
def rc(n, dt=np.complex64): return (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(dt)
dp = lambda v: v._fn("bdsp_device_ptr")(v._h)
# FFT sizes through every path (single vectors and small batches)
for n in (8, 1000, 1001, 4096, 8192, 16384, 1 << 15, 1 << 16, 1 << 17, 1 << 18, 1 << 19, 1 << 20, 1 << 21, 3 * 4096, 3 * (1 << 16)):
    DspVec(rc(n)).fft().ifft().to_numpy()
for n in (1000, 4096, 1 << 15, 3 * 4096):
    DspVec(rc(n, np.complex128)).fft().ifft().to_numpy()
for n, rows in ((64, 128), (128, 64), (256, 32), (512, 16), (1024, 8), (2048, 4), (4096, 3), (8192, 3), (16384, 2), (1 << 15, 2), (1 << 16, 2), (1 << 17, 1),
                (1 << 18, 1), (1 << 19, 1), (1 << 20, 1), (1 << 21, 1), (1 << 22, 1), (1 << 23, 1), (1 << 24, 1)):
    x = DspVec(rc(n * rows)); out = DspVec.zeros(2 * n * rows, is_complex=True, dtype=np.float32)
    for flags in (0, bd.F_SHIFT | bd.F_MAGNITUDE, bd.F_INVERSE, bd.F_INVERSE | bd.F_SHIFT):
        assert L.bdsp_fft_rows_c32(dp(x), dp(out), n, rows, flags) == 0
# real-input rows
for n, rows in ((256, 256), (1024, 64), (4096, 16), (16384, 4), (1 << 15, 2), (1 << 19, 1), (1 << 21, 1)):
    x = DspVec(rng.uniform(-1, 1, n * rows).astype(np.float32)); out = DspVec.zeros(2 * n * rows, is_complex=True, dtype=np.float32)
    for flags in (bd.F_REAL_INPUT, bd.F_REAL_INPUT | bd.F_SHIFT | bd.F_MAGNITUDE):
        assert L.bdsp_fft_rows_c32(dp(x), dp(out), n, rows, flags) == 0
# chirp-z and q*2^k batches
for n, rows in ((1000, 64), (3072, 16), (9973, 4)):
    x = DspVec(rc(n * rows)); out = DspVec.zeros(2 * n * rows, is_complex=True, dtype=np.float32)
    assert L.bdsp_fft_rows_c32(dp(x), dp(out), n, rows, 0) == 0
# convolution
for n, l in ((20000, 1023), (5000, 63), (9000, 2000), (40000, 5001), (300, 25)):
    DspVec(rc(n)).convolve_signal(DspVec(rc(l))).to_numpy()
DspVec(rc(20000)).convolve(bd.RAISED_COSINE, 0.35, 0.25, 31).to_numpy()
DspVec(rng.uniform(-1, 1, 20000).astype(np.float32)).convolve_signal(DspVec(rng.uniform(-1, 1, 200).astype(np.float32))).to_numpy()
# interpolation
r = rng.uniform(-1, 1, 20001).astype(np.float32)
DspVec(r).interpolatef(bd.SINC, 0.0, 4.0, 0.0, 12).to_numpy()
DspVec(r).interpolatef(bd.SINC, 0.0, 2.5, 0.1, 12).to_numpy()
DspVec(rc(5000)).interpolatef(bd.RAISED_COSINE, 0.35, 3.0, 0.0, 10).to_numpy()
DspVec(r).interpolate_lin(4.0, 0.0).to_numpy()
DspVec(r).interpolate_hermite(3.0, 0.0).to_numpy()
DspVec(r).interpolatei(bd.SINC, 0.0, 3).to_numpy()
DspVec(r).interpolate(bd.SINC, 0.0, 30011, 0.5).to_numpy()
DspVec(r).interpft(7001).to_numpy()
# symmetric transforms, windows, correlation
DspVec(r).plain_sfft().plain_sifft().to_numpy()
DspVec(r).windowed_sfft(bd.HAMMING).to_numpy()
DspVec(rc(3001)).apply_window(bd.BLACKMAN_HARRIS).windowed_fft(bd.HAMMING).to_numpy()
a, b = DspVec(rc(3001)), DspVec(rc(3001))
a.correlate(b.prepare_argument_padded()) if False else None
# elementwise math, reorganisation, reductions
v = DspVec(r)
for name in ("sin", "sqrt", "abs", "exp", "diff", "diff_with_start", "cum_sum"):
    DspVec(np.abs(r) + 0.1).math(name).to_numpy()
DspVec(r).math("unwrap", 6.28).to_numpy()
c = DspVec(rc(20001))
for name in ("sin", "sqrt", "ln", "tanh", "cum_sum", "diff"):
    DspVec(rc(20001)).math(name).to_numpy()
print(v.sum(), v.statistics()["max_index"], c.statistics()["min_index"], v.dot_product(DspVec(r)), c.statistics_split(3)[2]["count"])
parts = [DspVec(np.zeros(2, dtype=np.float32)) for _ in range(3)]
DspVec(r[:20001 - 20001 % 3]).split_into(parts)
DspVec(np.zeros(2, dtype=np.float32)).merge(parts).to_numpy()
DspVec(r).add_smaller(DspVec(r[:3])).to_numpy() if 20001 % 3 == 0 else None
# ---- round 2 kernels ----
# c64 tile passes (two / three passes, radix-3 first pass, real input, shifted, magnitude), generic fallbacks
for n in (1 << 14, 1 << 15, 1 << 17, 1 << 19, 1 << 20, 3 * (1 << 14), 3 * (1 << 17), 5 * (1 << 14)):
    DspVec(rc(n, np.complex128)).fft().ifft().to_numpy()
DspVec(rng.uniform(-1, 1, 1 << 15)).fft().to_numpy()
DspVec(rc(1 << 16, np.complex128)).fft_magnitude().to_numpy()
# fused c64 overlap-save blocks, 8192-point c32 blocks, tap counts at the block limits
for n, l in ((4096, 2), (20001, 1024), (70000, 513)):
    DspVec(rc(n, np.complex128)).convolve_signal(DspVec(rc(l, np.complex128))).to_numpy()
for n, l in ((8192, 4094), (100000, 3333), (1 << 16, 8191), (1 << 16, 4095)):
    DspVec(rc(n)).convolve_signal(DspVec(rc(l))).to_numpy()
# first-load multipliers: windows on every transform path, spectrum multiply of the full-length convolution, chirp filter
for n in (1000, 1001, 4096, 1 << 15, 1 << 16, 3 * (1 << 15), 1 << 21):
    DspVec(rc(n)).windowed_fft(bd.HAMMING).to_numpy()
for n in (1001, 1 << 14, 3 * (1 << 14)):
    DspVec(rc(n, np.complex128)).windowed_fft(bd.BLACKMAN_HARRIS).to_numpy()
DspVec(rc(1 << 15)).convolve_signal(DspVec(rc(9000))).to_numpy()
DspVec(rc(5 * (1 << 13))).convolve_signal(DspVec(rc(8500))).to_numpy()
DspVec(rc(30001, np.complex128)).convolve_signal(DspVec(rc(4100, np.complex128))).to_numpy()
# 16384-point rows (rolling pipeline), batched rows entry points
x = DspVec(rc(16384 * 5)); out = DspVec.zeros(2 * 16384 * 5, is_complex=True, dtype=np.float32)
for flags in (0, bd.F_SHIFT | bd.F_MAGNITUDE, bd.F_INVERSE):
    assert L.bdsp_fft_rows_c32(dp(x), dp(out), 16384, 5, flags) == 0
# later round-2 additions: windowed rows (separate pass + packed kernel, fused paths), c64 rows of 1024..4096 points, cluster
# transform for few 2^16-point rows, long c64 / c32 responses, correlation with the fused spectrum product
for n, rows, dt in ((1024, 64, np.complex64), (16384, 4, np.complex64), (1000, 8, np.complex64), (4096, 8, np.complex128), (1 << 15, 2, np.complex64)):
    xx = DspVec(rc(n * rows, dt)); oo = DspVec.zeros(2 * n * rows, is_complex=True, dtype=np.float32 if dt == np.complex64 else np.float64)
    fn = L.bdsp_fft_rows_c32 if dt == np.complex64 else L.bdsp_fft_rows_c64
    assert fn(dp(xx), dp(oo), n, rows, bd.F_SHIFT | bd.F_WINDOW(bd.HAMMING)) == 0
for n in (1024, 2048, 4096):
    xx = DspVec(rc(n * 6, np.complex128)); oo = DspVec.zeros(2 * n * 6, is_complex=True, dtype=np.float64)
    for flags in (0, bd.F_SHIFT | bd.F_MAGNITUDE, bd.F_INVERSE | bd.F_SHIFT):
        assert L.bdsp_fft_rows_c64(dp(xx), dp(oo), n, 6, flags) == 0
xx = DspVec(rc(65536 * 3)); oo = DspVec.zeros(2 * 65536 * 3, is_complex=True, dtype=np.float32)
for flags in (0, bd.F_SHIFT, bd.F_INVERSE | bd.F_SHIFT):
    assert L.bdsp_fft_rows_c32(dp(xx), dp(oo), 65536, 3, flags) == 0
DspVec(rc(259779, np.complex128)).convolve_signal(DspVec(rc(167, np.complex128))).to_numpy()
DspVec(rc(50000, np.complex128)).convolve_signal(DspVec(rc(2047, np.complex128))).to_numpy()
DspVec(rc(1 << 16)).convolve_signal(DspVec(rc(6000))).to_numpy()
a, b = DspVec(rc(3001)), DspVec(rc(3001))
a.correlate(b.prepare_argument_padded())
a.to_numpy()
L.bdsp_sync()
print("sanitize smoke done")
