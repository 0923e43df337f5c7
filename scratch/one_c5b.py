import sys
import numpy as np
sys.path.insert(0, ".")
import basic_dsp_b200 as bd
from basic_dsp_b200 import DspVec
n = 3 * (1 << 26)
v = DspVec.zeros(2 * n, is_complex=True, dtype=np.float64, init=0.5)
w = DspVec.zeros(2 * n, is_complex=True, dtype=np.float64, init=0.25)
mag = DspVec.zeros(n, is_complex=False, dtype=np.float64)
ph = DspVec.zeros(n, is_complex=False, dtype=np.float64)
for _ in range(2):
    v.scale_mul_mag_phase(complex(0.5, 0.25), w, mag, ph)
bd.lib().bdsp_sync()
