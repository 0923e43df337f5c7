import sys
import numpy as np
sys.path.insert(0, ".")
import basic_dsp_b200 as bd
from basic_dsp_b200 import DspVec
L = bd.lib()
rng = np.random.default_rng(0)
dp = lambda v: v._fn("bdsp_device_ptr")(v._h)
# fused c64 overlap-save block: 32 x 2^20, 1023 taps
n, rows = 1 << 20, 32
x = DspVec((rng.uniform(-1, 1, n * rows) + 1j * rng.uniform(-1, 1, n * rows)).astype(np.complex128))
out = DspVec.zeros(2 * n * rows, is_complex=True, dtype=np.float64)
h = DspVec((rng.uniform(-1, 1, 1023) + 1j * rng.uniform(-1, 1, 1023)).astype(np.complex128))
plan = L.bdsp_conv_plan_create_c64(dp(h), 1023)
for _ in range(2):
    L.bdsp_convolve_signal_rows_c64(dp(x), dp(out), n, rows, plan)
# c64 rows of 4096 points
for _ in range(2):
    L.bdsp_fft_rows_c64(dp(x), dp(out), 4096, n * rows // 4096, 0)
# single 2^16 c32 vector on the cluster kernel
v = DspVec((rng.uniform(-1, 1, 1 << 16) + 1j * rng.uniform(-1, 1, 1 << 16)).astype(np.complex64))
for _ in range(3):
    v.fft(); v.ifft()
L.bdsp_sync()
