"""Development aid: the TMA-staged kernel (bm_sad4.cu) against bm_sad3.cu on the same inputs (bit-exact raw disparity /
cost and final maps) over a few geometries, then the SAD/WTA stage time of both at the bench's batch size."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rt-depth-map_b200"))
import numpy as np, torch
import rtdm_b200 as rt
from rtdm_b200 import synth

def setk(k):
    if k == "4": os.environ.pop("RTDM_BM_KERNEL", None)
    else: os.environ["RTDM_BM_KERNEL"] = k

def run(W, H, nd, bs, seed, uniq=10, tex=10, roi=None, cap=31):
    L, R, _ = synth.stereo_pair(W, H, nd, seed)
    out = {}
    for k in ("3", "4"):
        setk(k)
        m = rt.CUDAMatcherKonolige(None, None, cap, bs, 0, tex, nd, nd, uniq, 100, 32, 1, max_width=W, max_height=H)
        if roi is not None: m.setROI1(roi)
        d = m.compute(L, R)
        out[k] = (d, m.debug_fetch(2, W, H), m.debug_fetch(3, W, H), m.last_kernel(), m.debug_fetch(1, W, H))
    a, b = out["3"], out["4"]
    h = bs // 2
    bad = int((a[1][h:H - h, nd - 1:] != b[1][h:H - h, nd - 1:]).sum()); badf = int((a[0] != b[0]).sum()); badr = int((a[4] != b[4]).sum())
    ok = a[1][h:H - h, nd - 1:] >= 0
    badc = int((a[2][h:H - h, nd - 1:][ok] != b[2][h:H - h, nd - 1:][ok]).sum())
    print(f"{W}x{H} nd={nd} bs={bs} uniq={uniq} tex={tex} roi={roi}: kernels {a[3]}/{b[3]} R' diff {badr} raw diff {bad} cost diff {badc} final diff {badf}", flush=True)
    if bad:
        ys, xs = np.nonzero(a[1][h:H - h, nd - 1:] != b[1][h:H - h, nd - 1:])
        print("   first diffs (y,x,k3,k4):", [(int(y) + h, int(x), int(a[1][y + h, x + nd - 1]), int(b[1][y + h, x + nd - 1])) for y, x in list(zip(ys, xs))[:8]])
        print("   x range", xs.min(), xs.max(), "y range", ys.min() + h, ys.max() + h, " x mod 2h hist", np.bincount(xs % (2 * h), minlength=2 * h))
    return bad + badc + badf

def timing(W, H, nd, bs, B):
    fr = [synth.stereo_pair(W, H, nd, 1000 + i) for i in range(4)]
    L = torch.from_numpy(np.stack([fr[i % 4][0] for i in range(B)])).cuda(); R = torch.from_numpy(np.stack([fr[i % 4][1] for i in range(B)])).cuda()
    D = torch.empty((B, H, W), dtype=torch.int16, device="cuda")
    st = torch.cuda.Stream()
    for k in ("3", "4"):
        setk(k)
        m = rt.CUDAMatcherKonolige(None, None, 31, bs, 0, 10, nd, nd, 10, 100, 32, 1, max_width=W, max_height=H, max_batch=B)
        go = lambda: m.compute_device(B, L.data_ptr(), W, W * H, R.data_ptr(), W, W * H, W, H, D.data_ptr(), W * 2, W * H * 2, st.cuda_stream)
        for _ in range(3): go()
        torch.cuda.synchronize(); m.set_profiling(True)
        for _ in range(10): go()
        t, c = m.stage_times()
        print(f"{W}x{H} nd={nd} bs={bs} B={B} kernel {m.last_kernel()}:", {s: round(v / c / B * 1e3, 2) for s, v in t.items()}, "us/frame", flush=True)

if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    tot = 0
    if what in ("all", "check"):
        tot += run(320, 240, 64, 13, 1)
        tot += run(1280, 720, 128, 13, 1000)
        tot += run(640, 480, 128, 9, 6)
        tot += run(640, 480, 64, 5, 7, uniq=0)
        tot += run(640, 480, 128, 15, 9)
        tot += run(333, 200, 64, 7, 10)
        tot += run(1280, 720, 128, 11, 1002)
        tot += run(401, 203, 64, 13, 8, tex=0)
        tot += run(934, 404, 192, 13, 11)
        tot += run(1280, 720, 256, 13, 12)
        tot += run(1280, 720, 32, 7, 13)
        tot += run(1280, 720, 128, 13, 1001, roi=(100, 50, 934, 404))
        print("TOTAL DIFF", tot, flush=True)
    if what in ("all", "time"):
        timing(1280, 720, 128, 13, 63)
        timing(1280, 720, 128, 15, 63)
        timing(1280, 720, 192, 13, 32)
        timing(1280, 720, 64, 9, 63)
        timing(934, 404, 192, 13, 1)
