"""One large frame, row bands over the GPUs of this process (rtdm_bm_rowband_*, peer copies) against one GPU:
bit-exactness and milliseconds per frame (host pointers, pinned, blocking calls).
  python tools/rowband_bench.py [n_gpus]"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rt-depth-map_b200"))
import numpy as np, torch
import rtdm_b200 as rt
from rtdm_b200 import synth


def run(W, H, nd, n, iters=20):
    L, R, _ = synth.stereo_pair(W, H, nd, 1000)
    Lp, Rp = torch.from_numpy(L).pin_memory().numpy(), torch.from_numpy(R).pin_memory().numpy()
    out = torch.empty((H, W), dtype=torch.int16).pin_memory().numpy()
    args = (None, None, 31, 13, 0, 10, nd, nd, 10, 100, 32, 1)
    one = rt.CUDAMatcherKonolige(*args, max_width=W, max_height=H)
    ref = one.compute(Lp, Rp).copy()
    for _ in range(3): one.compute(Lp, Rp, out)
    t0 = time.perf_counter()
    for _ in range(iters): one.compute(Lp, Rp, out)
    ms1 = (time.perf_counter() - t0) / iters * 1e3
    res = {"frame": f"{W}x{H}x{nd}", "one_gpu_ms": round(ms1, 3)}
    for k in sorted({2, 4, n} & set(range(2, n + 1))):
        rb = rt.CUDARowBandMatcherKonolige(*args, devices=list(range(k)), max_width=W, max_height=H)
        got = rb.compute(Lp, Rp).copy()
        for _ in range(3): rb.compute(Lp, Rp, out)
        t0 = time.perf_counter()
        for _ in range(iters): rb.compute(Lp, Rp, out)
        res[f"bands_{k}_gpus_ms"] = round((time.perf_counter() - t0) / iters * 1e3, 3)
        res[f"bands_{k}_bit_exact"] = bool(np.array_equal(got, ref))
        del rb
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else rt.device_count()
    run(3840, 2160, 256, n)
    run(1920, 1080, 128, n)
    run(1280, 720, 128, n)
