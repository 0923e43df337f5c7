"""Times bdsp_fft_rows_c32 for a few (points, rows) shapes of 2^26 points in total (512 MiB in, 512 MiB out)."""
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
total = 1 << 26
x = (rng.uniform(-1, 1, total) + 1j * rng.uniform(-1, 1, total)).astype(np.complex64)
vin = DspVec(x)
out = DspVec.zeros(2 * total, is_complex=True, dtype=np.float32)
for n in (64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 1 << 15, 1 << 16, 1 << 17, 1 << 18, 1 << 19, 1 << 20, 1 << 21, 1 << 22, 1 << 23, 1 << 24):
    med, best = T.run(lambda: L.bdsp_fft_rows_c32(dptr(vin), dptr(out), n, total // n, 0), 10)
    print("n=%8d rows=%6d  %.3f ms  %.0f GB/s (16 B/point)" % (n, total // n, med, 16 * total / med / 1e6))
