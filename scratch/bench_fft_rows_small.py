import sys
import numpy as np
sys.path.insert(0, ".")
import basic_dsp_b200 as bd
from basic_dsp_b200 import DspVec
from bench_configs import Timer, dptr
L = bd.lib(); bd.require_device(); T = Timer(L)
rng = np.random.default_rng(0)
total = 1 << 26
vin = DspVec((rng.uniform(-1, 1, total) + 1j * rng.uniform(-1, 1, total)).astype(np.complex64))
out = DspVec.zeros(2 * total, is_complex=True, dtype=np.float32)
for n in (16, 32, 64, 128, 100, 1000, 3 * 1024):
    rows = (total // n) // 16 * 16
    med, best = T.run(lambda: L.bdsp_fft_rows_c32(dptr(vin), dptr(out), n, rows, 0), 5)
    print("n=%6d rows=%8d  %.3f ms  %.0f GB/s (16 B/point)" % (n, rows, med, 16 * n * rows / med / 1e6))
