import sys, math
import numpy as np
sys.path.insert(0, ".")
import basic_dsp_b200 as bd
from basic_dsp_b200 import DspVec
from oracle import dsp_oracle as o
rng = np.random.default_rng(1)
bad = 0
for n in [3 * (1 << 14), 3 * (1 << 16), 15 * (1 << 14), 5 * (1 << 15), 7 * (1 << 17), 31 * (1 << 14), 9 * (1 << 20), 3 * (1 << 22), 17 * (1 << 18)]:
    x = (rng.uniform(-10, 10, n) + 1j * rng.uniform(-10, 10, n)).astype(np.complex128)
    v = DspVec(x)
    X = v.plain_fft().to_numpy()
    e = o.rel_l2(X, o.plain_fft(x))
    back = v.plain_ifft().to_numpy() / n
    e2 = o.rel_l2(back, x)
    Xs = DspVec(x).fft().to_numpy()
    e3 = o.rel_l2(Xs, o.fft(x))
    ok = e <= 1e-12 * math.log2(n) and e2 <= 2e-12 * math.log2(n) and e3 <= 1e-12 * math.log2(n)
    bad += not ok
    print("n=%10d  fwd %.2e  roundtrip %.2e  shifted %.2e  %s" % (n, e, e2, e3, "ok" if ok else "FAIL"), flush=True)
sys.exit(1 if bad else 0)
