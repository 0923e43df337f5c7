"""quick parity + timing of convolve_signal rows for a few (N, L, rows): fused 4096 / 8192 kernels vs numpy."""
import sys
import numpy as np
sys.path.insert(0, ".")
import basic_dsp_b200 as bd
from basic_dsp_b200 import DspVec
from bench_configs import Timer, dptr
from oracle import dsp_oracle as o

L = bd.lib()
bd.require_device()
T = Timer(L)
rng = np.random.default_rng(0)
bad = 0
for n, taps, rows in [(8192, 1023, 1), (8192, 4094, 2), (20000, 1023, 40), (20001, 1022, 40), (1 << 16, 2047, 12), (1 << 16, 3000, 12),
                      (1 << 16, 4094, 12), (1 << 17, 63, 8), (12345, 2, 64), (1 << 20, 1023, 2)]:
    x = (rng.uniform(-10, 10, n * rows) + 1j * rng.uniform(-10, 10, n * rows)).astype(np.complex64)
    h = ((rng.uniform(-1, 1, taps) + 1j * rng.uniform(-1, 1, taps)) / 10).astype(np.complex64)
    xv, hv, out = DspVec(x), DspVec(h), DspVec.zeros(2 * n * rows, is_complex=True)
    plan = L.bdsp_conv_plan_create_c32(dptr(hv), taps)
    rc = L.bdsp_convolve_signal_rows_c32(dptr(xv), dptr(out), n, rows, plan)
    got = out.to_numpy().reshape(rows, n)
    worst = max(o.rel_l2(got[r], o.convolve_signal(x.reshape(rows, n)[r], h)) for r in range(rows))
    ok = rc == 0 and worst <= 1.2e-4
    bad += not ok
    print("N=%8d L=%5d rows=%3d rc=%d rel_l2=%.2e %s" % (n, taps, rows, rc, worst, "ok" if ok else "FAIL"), flush=True)
    L.bdsp_conv_plan_destroy(plan)
n, rows = 1 << 20, 64
x = (rng.uniform(-1, 1, n * rows) + 1j * rng.uniform(-1, 1, n * rows)).astype(np.complex64)
vin = DspVec(x)
out = DspVec.zeros(2 * n * rows, is_complex=True, dtype=np.float32)
for taps in (31, 255, 1023, 2047, 3071, 4094):
    h = (rng.uniform(-1, 1, taps) + 1j * rng.uniform(-1, 1, taps)).astype(np.complex64)
    hv = DspVec(h)
    plan = L.bdsp_conv_plan_create_c32(dptr(hv), taps)
    med, best = T.run(lambda: L.bdsp_convolve_signal_rows_c32(dptr(vin), dptr(out), n, rows, plan), 10)
    print("taps=%5d  %.4f ms  %.0f GB/s (16 B/sample)  frac %.3f" % (taps, med, 16 * n * rows / med / 1e6, 16 * n * rows / med / 1e6 / 6548.5))
    L.bdsp_conv_plan_destroy(plan)
sys.exit(1 if bad else 0)
