"""Times bdsp_fft_rows_c64 for a few (points, rows) shapes of 2^25 points in total (512 MiB in, 512 MiB out)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import basic_dsp_b200 as bd
from basic_dsp_b200 import DspVec
from bench_configs import Timer, dptr

L = bd.lib()
bd.require_device()
T = Timer(L)
rng = np.random.default_rng(0)
total = 1 << 25
x = (rng.uniform(-1, 1, total) + 1j * rng.uniform(-1, 1, total)).astype(np.complex128)
vin = DspVec(x)
out = DspVec.zeros(2 * total, is_complex=True, dtype=np.float64)
for n in (256, 1024, 4096, 8192, 1 << 14, 1 << 16, 1 << 18, 1 << 20, 1 << 22):
    med, best = T.run(lambda: L.bdsp_fft_rows_c64(dptr(vin), dptr(out), n, total // n, 0), 10)
    print("n=%8d rows=%6d  %.3f ms  %.0f GB/s (32 B/point)" % (n, total // n, med, 32 * total / med / 1e6))
