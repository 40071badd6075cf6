"""Development aid: the warp-specialised kernel (bm_sad3.cu) against the bm_sad2.cu kernel on the same inputs
(bit-exact raw disparity / cost and final maps), over a few geometries, then per-stage timing of both."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rt-depth-map_b200"))
import numpy as np, torch
import rtdm_b200 as rt
from rtdm_b200 import synth

def run(W, H, nd, bs, seed, uniq=10, tex=10, roi=None):
    L, R, _ = synth.stereo_pair(W, H, nd, seed)
    out = {}
    for k in ("2", "3"):
        if k == "2": os.environ["RTDM_BM_KERNEL"] = "2"
        else: os.environ.pop("RTDM_BM_KERNEL", None)
        m = rt.CUDAMatcherKonolige(None, None, 31, bs, 0, tex, nd, nd, uniq, 100, 32, 1, max_width=W, max_height=H)
        if roi is not None: m.setROI1(roi)
        d = m.compute(L, R)
        out[k] = (d, m.debug_fetch(2, W, H), m.debug_fetch(3, W, H), m.last_kernel())
    a, b = out["2"], out["3"]
    bad = int((a[1] != b[1]).sum()); badc = int((a[2] != b[2]).sum()); badf = int((a[0] != b[0]).sum())
    print(f"{W}x{H} nd={nd} bs={bs} uniq={uniq} tex={tex} roi={roi}: kernels {a[3]}/{b[3]} raw diff {bad} cost diff {badc} final diff {badf}", flush=True)
    if bad:
        ys, xs = np.nonzero(a[1] != b[1])
        print("   first diffs (y,x,k2,k3):", [(int(y), int(x), int(a[1][y, x]), int(b[1][y, x])) for y, x in list(zip(ys, xs))[:8]])
        print("   x range", xs.min(), xs.max(), "y range", ys.min(), ys.max(), " x mod 12 hist", np.bincount((xs - (nd - 1)) % 12, minlength=12))
    return bad + badc + badf

tot = 0
tot += run(320, 240, 64, 13, 1)
tot += run(1280, 720, 128, 13, 1000)
tot += run(640, 480, 128, 13, 5)
tot += run(640, 480, 128, 9, 6)
tot += run(640, 480, 64, 5, 7, uniq=0)
tot += run(640, 480, 128, 15, 9)
tot += run(333, 200, 64, 7, 10)
tot += run(1280, 720, 128, 11, 1002)
tot += run(401, 203, 64, 13, 8, tex=0)
tot += run(1280, 720, 128, 13, 1001, roi=(100, 50, 934, 404))
print("TOTAL DIFF", tot)
