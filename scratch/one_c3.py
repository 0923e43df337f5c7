import sys
import numpy as np
sys.path.insert(0, ".")
import basic_dsp_b200 as bd
from basic_dsp_b200 import DspVec
L = bd.lib()
n, rows = 1 << 14, 4096
rng = np.random.default_rng(0)
x = DspVec((rng.uniform(-1, 1, n * rows) + 1j * rng.uniform(-1, 1, n * rows)).astype(np.complex64))
out = DspVec.zeros(n * rows, is_complex=False, dtype=np.float32)
dp = lambda v: v._fn("bdsp_device_ptr")(v._h)
for _ in range(3):
    assert L.bdsp_fft_rows_c32(dp(x), dp(out), n, rows, bd.F_SHIFT | bd.F_MAGNITUDE) == 0
L.bdsp_sync()
