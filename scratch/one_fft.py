import sys
import numpy as np
sys.path.insert(0, ".")
import basic_dsp_b200 as bd
from basic_dsp_b200 import DspVec
from bench_configs import dptr
n = int(sys.argv[1]); total = 1 << 26
L = bd.lib(); bd.require_device()
x = np.ones(total, dtype=np.complex64)
vin = DspVec(x); out = DspVec.zeros(2 * total, is_complex=True, dtype=np.float32)
for _ in range(3):
    L.bdsp_fft_rows_c32(dptr(vin), dptr(out), n, total // n, 0)
L.bdsp_sync()
