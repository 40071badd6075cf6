"""Development aid: soak test of the warp-specialised kernel (bm_sad3.cu) against bm_sad2.cu over random geometries
(stripe and band boundaries, ROIs, caps, thresholds), raw WTA output + cost + final map, batches of 1..5."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rt-depth-map_b200"))
import numpy as np
import rtdm_b200 as rt
from rtdm_b200 import synth
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
N = int(sys.argv[2]) if len(sys.argv) > 2 else 120
bad = 0
for t in range(N):
    nd = int(rng.choice([32, 48, 64, 96, 128, 192, 256])); bs = int(rng.choice([5, 7, 9, 11, 13, 15]))
    W = int(rng.integers(nd + bs + 2, 1500)); H = int(rng.integers(bs + 2, 260))
    if t % 7 == 0: W = nd - 1 + int(rng.choice([165, 166, 180, 181, 330, 331, 1153]))     # stripe-width boundaries
    B = int(rng.integers(1, 6))
    cap, tex, uniq, d12 = int(rng.integers(1, 32)), int(rng.integers(0, 60)), int(rng.integers(0, 30)), int(rng.integers(-1, 3))
    roi = None
    if t % 3 == 0:
        x0, y0 = int(rng.integers(0, W // 2)), int(rng.integers(0, H // 2))
        roi = (x0, y0, int(rng.integers(1, W - x0 + 1)), int(rng.integers(1, H - y0 + 1)))
    fr = [synth.stereo_pair(W, H, nd, 9000 + 10 * t + i) for i in range(B)]
    L = np.stack([f[0] for f in fr]); R = np.stack([f[1] for f in fr])
    res = {}
    for k in ("2", "3"):
        if k == "2": os.environ["RTDM_BM_KERNEL"] = "2"
        else: os.environ.pop("RTDM_BM_KERNEL", None)
        m = rt.CUDAMatcherKonolige(None, None, cap, bs, 0, tex, nd, nd, uniq, 100, 32, d12, max_width=W, max_height=H, max_batch=B)
        if roi is not None: m.setROI1(roi)
        d = m.compute_batch(L, R)
        res[k] = (d, m.debug_fetch(2, W, H), m.debug_fetch(3, W, H), m.last_kernel())
    a, b = res["2"], res["3"]
    ok = np.array_equal(a[0], b[0])
    if ok and roi is None and b[3] == 3:
        # raw WTA output and cost of the last chunk's first frame, inside the computed rows / columns
        h = bs // 2
        ra, rb = a[1][h:H - h, nd - 1:], b[1][h:H - h, nd - 1:]
        ok = np.array_equal(ra, rb) and np.array_equal(a[2][h:H - h, nd - 1:][ra >= 0], b[2][h:H - h, nd - 1:][rb >= 0])
    if not ok or t % 20 == 0:
        print(t, W, H, nd, bs, B, cap, tex, uniq, d12, roi, "kernels", a[3], b[3], "OK" if ok else "MISMATCH %d" % int((a[0] != b[0]).sum()), flush=True)
    bad += not ok
print("soak done:", N, "cases,", bad, "mismatches")
