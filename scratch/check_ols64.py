import sys, math
import numpy as np
sys.path.insert(0, ".")
import basic_dsp_b200 as bd
from basic_dsp_b200 import DspVec
from oracle import dsp_oracle as o
rng = np.random.default_rng(3)
bad = 0
for n, l in [(4096, 1), (4096, 2), (5000, 25), (5000, 63), (10000, 255), (20000, 1023), (20001, 1024), (20001, 1025), (1 << 16, 777), (4095, 100), (3 * 4096 + 5, 1000)]:
    x = (rng.uniform(-10, 10, n) + 1j * rng.uniform(-10, 10, n)).astype(np.complex128)
    h = ((rng.uniform(-1, 1, l) + 1j * rng.uniform(-1, 1, l))).astype(np.complex128)
    got = DspVec(x).convolve_signal(DspVec(h)).to_numpy()
    ref = o.convolve_signal_direct(x, h) if n * l < 3e7 else o.convolve_signal(x, h)
    e = o.rel_l2(got, ref)
    ok = e <= 1e-12 * 12
    bad += not ok
    print("n=%7d l=%5d  err %.2e %s" % (n, l, e, "ok" if ok else "FAIL"), flush=True)
sys.exit(1 if bad else 0)
