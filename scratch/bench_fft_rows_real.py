"""bdsp_fft_rows_c32 with BDSP_F_REAL_INPUT: 2^26 real f32 points in total (256 MiB in, 512 MiB out)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import basic_dsp_b200 as bd
from basic_dsp_b200 import DspVec
from bench_configs import Timer, dptr

L = bd.lib(); bd.require_device(); T = Timer(L)
rng = np.random.default_rng(0)
total = 1 << 26
vin = DspVec(rng.uniform(-1, 1, total).astype(np.float32))
out = DspVec.zeros(2 * total, is_complex=True, dtype=np.float32)
for n in (64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 1 << 16, 1 << 20):
    med, best = T.run(lambda: L.bdsp_fft_rows_c32(dptr(vin), dptr(out), n, total // n, bd.F_REAL_INPUT), 10)
    print("n=%8d rows=%6d  %.3f ms  %.0f GB/s (12 B/point)" % (n, total // n, med, 12 * total / med / 1e6))
