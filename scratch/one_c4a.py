import sys
import numpy as np
sys.path.insert(0, ".")
import basic_dsp_b200 as bd
from basic_dsp_b200 import DspVec
n = 1 << 24
rng = np.random.default_rng(0)
x = rng.uniform(-1, 1, n).astype(np.float32)
v = DspVec(x)
for _ in range(3):
    v.set_len(n)
    v.interpolatef(bd.SINC, 0.0, 4.0, 0.0, 12)
bd.lib().bdsp_sync()
