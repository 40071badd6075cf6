"""Development aid: where the end-to-end (host pointers, pinned) BM + filter pipeline loses time.
Times the matcher stream alone, the filter alone, and both, per 64-frame batch."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rt-depth-map_b200"))
import numpy as np, torch
import rtdm_b200 as rt
from rtdm_b200 import synth
W, H, ND, B = 1280, 720, 128, int(sys.argv[1]) if len(sys.argv) > 1 else 64
fr = [synth.stereo_pair(W, H, ND, 1000 + i) for i in range(4)]
pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
L = pin(np.stack([fr[i % 4][0] for i in range(B)])); R = pin(np.stack([fr[i % 4][1] for i in range(B)]))
M = pin(np.stack([synth.binary_mask(W, H, 3000 + (i % 4)) for i in range(B)]))
D = [torch.empty((B, H, W), dtype=torch.int16).pin_memory().numpy() for _ in range(2)]
MO = torch.empty((B, H, W), dtype=torch.uint8).pin_memory().numpy()
m = rt.CUDAMatcherKonolige(None, None, 31, 13, 0, 10, ND, ND, 10, 100, 32, 1, max_width=W, max_height=H, max_batch=B)
f = rt.CUDAMorphologicalFilter(W, H, 8, max_batch=B)
def run(matcher, filt, steps=8):
    def step(i):
        if matcher: m.submit_batch(L, R, D[i & 1])
        if filt: f.run_batch_async(M, MO)
        if matcher and i > 0: m.wait_oldest()
        if filt: f.sync()
    step(0)
    if matcher: m.wait()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for i in range(steps): step(i)
    if matcher: m.wait()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / steps * 1e3
def run2(steps=8):
    """filter result read one submission later, like the matcher's"""
    def step(i):
        m.submit_batch(L, R, D[i & 1])
        if i > 0: f.sync()
        f.run_batch_async(M, MO)
        if i > 0: m.wait_oldest()
    step(0); m.wait(); f.sync(); torch.cuda.synchronize()
    t = time.perf_counter()
    for i in range(steps): step(i)
    m.wait(); f.sync(); torch.cuda.synchronize()
    return (time.perf_counter() - t) / steps * 1e3
print("chunk", os.environ.get("RTDM_BM_CHUNK", "default"), "B", B)
print("matcher only  ms/batch", round(run(True, False), 3))
print("filter only   ms/batch", round(run(False, True), 3))
print("both          ms/batch", round(run(True, True), 3), "->", round(B / run(True, True) , 2), "kfps")
r2 = run2()
print("both, filter read one submission later ms/batch", round(r2, 3), "->", round(B / r2, 2), "kfps")
# matcher blocking
t = time.perf_counter()
for i in range(4): m.compute_batch(L, R, D[0])
print("matcher blocking ms/batch", round((time.perf_counter() - t) / 4 * 1e3, 3))
