"""Single-frame latency through the host plugin API (what Estimator::run would see per frame)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rt-depth-map_b200"))
import numpy as np, torch
import rtdm_b200 as rt
from rtdm_b200 import synth

def timeit(fn, n=30):
    for _ in range(5): fn()
    t = time.perf_counter()
    for _ in range(n): fn()
    return (time.perf_counter() - t) / n * 1e3

for (W, H, nd, tag) in [(1280, 720, 128, "720p full frame"), (934, 404, 128, "720p calibrated ROI crop"), (640, 480, 128, "640x480")]:
    L, R, _ = synth.stereo_pair(W, H, nd, 1000)
    Lp, Rp = torch.from_numpy(L).pin_memory().numpy(), torch.from_numpy(R).pin_memory().numpy()
    out = torch.empty((H, W), dtype=torch.int16).pin_memory().numpy()
    bm = rt.CUDAMatcherKonolige(None, None, 31, 13, 0, 10, nd, nd, 10, 100, 32, 1, max_width=W, max_height=H)
    ms_bm = timeit(lambda: bm.compute(Lp, Rp, out))
    sg = rt.CUDASemiGlobalMatcher(5, 0, nd, 10, 100, 32, 1, mode=0, max_width=W, max_height=H)
    ms_sg = timeit(lambda: sg.compute(Lp, Rp, out), 10)
    hh = rt.CUDASemiGlobalMatcher(5, 0, nd, 10, 100, 32, 1, mode=1, max_width=W, max_height=H)
    ms_hh = timeit(lambda: hh.compute(Lp, Rp, out), 10)
    f = rt.CUDAMorphologicalFilter(W, H, 8)
    f.getVideoInBuffer()[:] = synth.binary_mask(W, H, 3)
    ms_f = timeit(lambda: f.run())
    print(f"{tag} {W}x{H} nd={nd}: BM compute {ms_bm:.3f} ms, SGBM {ms_sg:.2f} ms, SGBM-HH {ms_hh:.2f} ms, filter run {ms_f:.3f} ms (host pointers, sync per call)")
    del bm, sg, hh, f

# mask front-end / back-end at the 720p calibrated ROI: 1280x720 RGB frame -> 934x404 mask -> boxes
W, H, roi = 1280, 720, (173, 158, 934, 404)
rng = np.random.default_rng(1)
ys, xs = np.mgrid[0:H, 0:W]
fx, fy = xs * 32 + ys + 11, ys * 32 + 7
m1 = np.stack([fx >> 5, fy >> 5], -1).astype(np.int16); m2 = (((fy & 31) << 5) | (fx & 31)).astype(np.uint16)
rgb = torch.from_numpy(rng.integers(0, 256, (H, W, 3)).astype(np.uint8)).pin_memory().numpy()
cm = rt.CUDAColorMask(m1, m2, roi)
ms_cm = timeit(lambda: cm.run(rgb, (30, 60, 50), (100, 255, 255)))
mask = synth.binary_mask(roi[2], roi[3], 3000)
reg = rt.CUDAObjectRegions(roi[2], roi[3], 1024)
ms_reg = timeit(lambda: reg.run(mask, 400))
dm = torch.from_numpy(mask).cuda()
ms_regd = timeit(lambda: reg.run_device(dm.data_ptr(), roi[2], roi[2], roi[3], 400))
print(f"mask front-end 1280x720 RGB -> {roi[2]}x{roi[3]} mask (host pointers): {ms_cm:.3f} ms; object boxes of a {roi[2]}x{roi[3]} mask: "
      f"{ms_reg:.3f} ms from host memory, {ms_regd:.3f} ms device-resident ({len(reg.run(mask, 400)[0])} boxes)")
