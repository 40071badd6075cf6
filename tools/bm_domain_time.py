import os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/rt-depth-map_b200')
import numpy as np, torch
import rtdm_b200 as rt
from rtdm_b200 import synth
for (W, H, nd, bs) in [(1280, 720, 192, 13), (640, 480, 96, 13), (320, 240, 48, 13), (1280, 720, 256, 13), (1280, 720, 32, 13), (1280, 720, 128, 15), (1280, 720, 64, 9)]:
    B = 16
    fr = [synth.stereo_pair(W, H, nd, 1000 + i) for i in range(2)]
    L = torch.from_numpy(np.stack([fr[i % 2][0] for i in range(B)])).cuda(); R = torch.from_numpy(np.stack([fr[i % 2][1] for i in range(B)])).cuda()
    D = torch.empty((B, H, W), dtype=torch.int16, device="cuda")
    st = torch.cuda.Stream()
    for k in ("2", "3"):
        if k == "2": os.environ["RTDM_BM_KERNEL"] = "2"
        else: os.environ.pop("RTDM_BM_KERNEL", None)
        m = rt.CUDAMatcherKonolige(None, None, 31, bs, 0, 10, nd, nd, 10, 100, 32, 1, max_width=W, max_height=H, max_batch=B)
        run = lambda: m.compute_device(B, L.data_ptr(), W, W * H, R.data_ptr(), W, W * H, W, H, D.data_ptr(), W * 2, W * H * 2, st.cuda_stream)
        for _ in range(3): run()
        torch.cuda.synchronize(); m.set_profiling(True)
        for _ in range(5): run()
        t, c = m.stage_times()
        print(W, H, nd, bs, "kernel", m.last_kernel(), "sad_wta us/frame", round(t["sad_wta"] / c / B * 1e3, 1), flush=True)
