"""Development aid: fused SGBM cost kernel against the two-kernel path on assorted shapes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rt-depth-map_b200"))
import numpy as np
import rtdm_b200 as rt
from rtdm_b200 import synth
for (W, H, nd, bs) in [(233, 157, 32, 3), (233, 157, 32, 5), (288, 157, 32, 3), (233, 157, 64, 3), (320, 240, 64, 5), (300, 100, 16, 1), (700, 100, 128, 7), (500, 60, 256, 5)]:
    L, R, _ = synth.stereo_pair(W, H, nd, 5)
    outs = []
    for old in (True, False):
        if old: os.environ["RTDM_SGBM_OLDCOST"] = "1"
        else: os.environ.pop("RTDM_SGBM_OLDCOST", None)
        m = rt.CUDASemiGlobalMatcher(bs, 0, nd, 10, 0, 0, 1, mode=0, max_width=W, max_height=H)
        outs.append(m.compute(L, R))
    bad = np.argwhere(outs[0] != outs[1])
    print(W, H, nd, bs, "mismatch", len(bad), (bad.min(0), bad.max(0)) if len(bad) else "")
