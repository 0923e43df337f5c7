import sys
import numpy as np
sys.path.insert(0, ".")
import basic_dsp_b200 as bd
from basic_dsp_b200 import DspVec
L = bd.lib()
rng = np.random.default_rng(0)
dp = lambda v: v._fn("bdsp_device_ptr")(v._h)
n, rows, taps = 1 << 20, 64, 2047
x = DspVec((rng.uniform(-1, 1, n * rows) + 1j * rng.uniform(-1, 1, n * rows)).astype(np.complex64))
out = DspVec.zeros(2 * n * rows, is_complex=True, dtype=np.float32)
h = DspVec((rng.uniform(-1, 1, taps) + 1j * rng.uniform(-1, 1, taps)).astype(np.complex64))
plan = L.bdsp_conv_plan_create_c32(dp(h), taps)
for _ in range(3):
    L.bdsp_convolve_signal_rows_c32(dp(x), dp(out), n, rows, plan)
L.bdsp_sync()
