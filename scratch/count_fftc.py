import sys
import numpy as np
sys.path.insert(0, ".")
import basic_dsp_b200 as bd
from basic_dsp_b200 import DspVec
n = 1 << 16
x = (np.ones(n) + 0j).astype(np.complex64)
v = DspVec(x)
v.fft(); v.ifft()
b = bd.kernel_launch_count()
v.fft()
print("fft:", bd.kernel_launch_count() - b)
b = bd.kernel_launch_count()
v.ifft()
print("ifft:", bd.kernel_launch_count() - b)
bd.lib().bdsp_sync()
