import sys, math
import numpy as np
sys.path.insert(0, ".")
import basic_dsp_b200 as bd
from basic_dsp_b200 import DspVec
from oracle import dsp_oracle as o
rng = np.random.default_rng(5)
n = 1 << 16
x = (rng.uniform(-10, 10, n) + 1j * rng.uniform(-10, 10, n)).astype(np.complex64)
before = bd.kernel_launch_count()
X = DspVec(x).plain_fft().to_numpy()
print("launches for one plain_fft:", bd.kernel_launch_count() - before)
print("plain_fft", o.rel_l2(X, o.plain_fft(x)))
print("fft", o.rel_l2(DspVec(x).fft().to_numpy(), o.fft(x)))
Xf = o.fft(x).astype(np.complex64)
print("ifft", o.rel_l2(DspVec(Xf, domain=bd.FREQ).ifft().to_numpy(), o.ifft(Xf)))
print("plain_ifft", o.rel_l2(DspVec(X, domain=bd.FREQ).plain_ifft().to_numpy() / n, x))
print("roundtrip", o.rel_l2(DspVec(x).fft().ifft().to_numpy(), x))
L = bd.lib()
dp = lambda v: v._fn("bdsp_device_ptr")(v._h)
rows = 3
xs = (rng.uniform(-1, 1, n * rows) + 1j * rng.uniform(-1, 1, n * rows)).astype(np.complex64)
vin = DspVec(xs); out = DspVec.zeros(2 * n * rows, is_complex=True, dtype=np.float32)
assert L.bdsp_fft_rows_c32(dp(vin), dp(out), n, rows, 0) == 0
got = out.to_numpy().reshape(rows, n)
print("rows", max(o.rel_l2(got[r], o.plain_fft(xs.reshape(rows, n)[r])) for r in range(rows)))
