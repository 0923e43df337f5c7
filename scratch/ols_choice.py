"""64 x 2^20 c32 convolve_signal rows: 4096- vs 8192-point fused blocks over the tap count (BDSP_OLS_FORCE)."""
import os, sys
import numpy as np
sys.path.insert(0, ".")
import basic_dsp_b200 as bd
from basic_dsp_b200 import DspVec
from bench_configs import Timer, dptr
L = bd.lib(); bd.require_device(); T = Timer(L)
rng = np.random.default_rng(0)
n, rows = 1 << 20, 64
x = (rng.uniform(-1, 1, n * rows) + 1j * rng.uniform(-1, 1, n * rows)).astype(np.complex64)
vin = DspVec(x); out = DspVec.zeros(2 * n * rows, is_complex=True, dtype=np.float32)
for taps in [int(a) for a in sys.argv[1:]]:
    h = (rng.uniform(-1, 1, taps) + 1j * rng.uniform(-1, 1, taps)).astype(np.complex64)
    hv = DspVec(h)
    plan = L.bdsp_conv_plan_create_c32(dptr(hv), taps)
    med, best = T.run(lambda: L.bdsp_convolve_signal_rows_c32(dptr(vin), dptr(out), n, rows, plan), 10)
    print("force=%s taps=%5d  %.4f ms  frac %.3f" % (os.environ.get("BDSP_OLS_FORCE", "-"), taps, med, 16 * n * rows / med / 1e6 / 6548.5), flush=True)
    L.bdsp_conv_plan_destroy(plan)
