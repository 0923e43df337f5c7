import sys
import numpy as np
sys.path.insert(0, ".")
import basic_dsp_b200 as bd
from basic_dsp_b200 import DspVec
n = 3 * (1 << 26)
v = DspVec.zeros(2 * n, is_complex=True, dtype=np.float64, init=0.5)
v.plain_fft(); v.plain_ifft(); v.plain_fft()
bd.lib().bdsp_sync()
