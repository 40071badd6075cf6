"""Development aid: does a second matcher on a second stream overlap its small kernels with the first one's SAD kernel?
One handle with 63 frames on one stream against two handles (35 + 28 frames) on two streams."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rt-depth-map_b200"))
import numpy as np, torch
import rtdm_b200 as rt
from rtdm_b200 import synth
W, H, nd = 1280, 720, 128
fr = [synth.stereo_pair(W, H, nd, 1000 + i) for i in range(4)]
def dev(n):
    L = torch.from_numpy(np.stack([fr[i % 4][0] for i in range(n)])).cuda(); R = torch.from_numpy(np.stack([fr[i % 4][1] for i in range(n)])).cuda()
    return L, R, torch.empty((n, H, W), dtype=torch.int16, device="cuda")
def mk(n): return rt.CUDAMatcherKonolige(None, None, 31, 13, 0, 10, nd, nd, 10, 100, 32, 1, max_width=W, max_height=H, max_batch=n)
def run(m, bufs, st, n):
    L, R, D = bufs
    m.compute_device(n, L.data_ptr(), W, W * H, R.data_ptr(), W, W * H, W, H, D.data_ptr(), W * 2, W * H * 2, st.cuda_stream)
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / reps * 1e3
m63, b63, s0 = mk(63), dev(63), torch.cuda.Stream()
print("one stream, 63 frames: %.3f ms" % timeit(lambda: run(m63, b63, s0, 63)))
for (na, nb) in [(35, 28), (42, 21), (32, 31)]:
    ma, mb, ba, bb = mk(na), mk(nb), dev(na), dev(nb)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    print("two streams, %d + %d frames: %.3f ms" % (na, nb, timeit(lambda: (run(ma, ba, s1, na), run(mb, bb, s2, nb)))))
    h1, h2 = torch.cuda.Stream(priority=-1), torch.cuda.Stream(priority=0)
    print("  first stream at high priority: %.3f ms" % timeit(lambda: (run(ma, ba, h1, na), run(mb, bb, h2, nb))))
    print("  same, one stream: %.3f ms" % timeit(lambda: (run(ma, ba, s1, na), run(mb, bb, s1, nb))))
    del ma, mb
