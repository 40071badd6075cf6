import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rt-depth-map_b200"))
import numpy as np
import rtdm_b200 as rt
from rtdm_b200 import synth
W, H, nd, bs = 640, 480, 64, 5
L, R, _ = synth.stereo_pair(W, H, nd, 7)
os.environ["RTDM_BM_KERNEL"] = "2"
m = rt.CUDAMatcherKonolige(None, None, 31, bs, 0, 10, nd, nd, 0, 100, 32, 1, max_width=W, max_height=H)
m.compute(L, R); ref = m.debug_fetch(2, W, H)
os.environ.pop("RTDM_BM_KERNEL")
for t in range(6):
    m.compute(L, R); got = m.debug_fetch(2, W, H)
    ys, xs = np.nonzero(ref != got)
    print(os.environ.get("RTDM_BM_DEBUG"), "try", t, "diff", len(ys), sorted(set(zip(ys.tolist(), xs.tolist())))[:10], flush=True)
