import os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/rt-depth-map_b200')
import numpy as np
import rtdm_b200 as rt
from rtdm_b200 import synth
W,H,nd,bs=320,240,64,13
L,R,_=synth.stereo_pair(W,H,nd,1)
m = rt.CUDAMatcherKonolige(None, None, 31, bs, 0, 10, nd, nd, 10, 100, 32, 1, max_width=W, max_height=H)
try:
    d=m.compute(L,R); print("kernel", m.last_kernel(), (d>=0).mean())
except Exception as e: print("ERR", e)
