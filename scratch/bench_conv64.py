"""convolve_signal rows in f64: 32 x 2^20 c64, 1023 taps (generic overlap-save kernel)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import basic_dsp_b200 as bd
from basic_dsp_b200 import DspVec
from bench_configs import Timer, dptr

L = bd.lib(); bd.require_device(); T = Timer(L)
rng = np.random.default_rng(0)
n, rows = 1 << 20, 32
x = (rng.uniform(-1, 1, n * rows) + 1j * rng.uniform(-1, 1, n * rows)).astype(np.complex128)
vin = DspVec(x); out = DspVec.zeros(2 * n * rows, is_complex=True, dtype=np.float64)
for taps in (255, 1023, 2047):
    h = (rng.uniform(-1, 1, taps) + 1j * rng.uniform(-1, 1, taps)).astype(np.complex128)
    hv = DspVec(h)
    plan = L.bdsp_conv_plan_create_c64(dptr(hv), taps)
    med, best = T.run(lambda: L.bdsp_convolve_signal_rows_c64(dptr(vin), dptr(out), n, rows, plan), 5)
    print("c64 taps=%5d  %.3f ms  %.0f GB/s (32 B/sample)" % (taps, med, 32 * n * rows / med / 1e6))
    L.bdsp_conv_plan_destroy(plan)
