"""Device-resident SGBM timing at other disparity counts / sizes (development aid):  python tools/quick_time_sgbm_nd.py W H nd batch"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rt-depth-map_b200"))
import numpy as np, torch
import rtdm_b200 as rt
from rtdm_b200 import synth
W, H, nd, B = (int(v) for v in sys.argv[1:5])
for mode in (0, 1):
    frames = [synth.stereo_pair(W, H, nd, 1000 + i) for i in range(2)]
    L = torch.from_numpy(np.stack([frames[i % 2][0] for i in range(B)])).cuda()
    R = torch.from_numpy(np.stack([frames[i % 2][1] for i in range(B)])).cuda()
    D = torch.empty((B, H, W), dtype=torch.int16, device="cuda")
    m = rt.CUDASemiGlobalMatcher(5, 0, nd, 10, 100, 32, 1, mode=mode, max_width=W, max_height=H, max_batch=B)
    st = torch.cuda.Stream()
    def run():
        m.compute_device(B, L.data_ptr(), W, W * H, R.data_ptr(), W, W * H, W, H, D.data_ptr(), W * 2, W * H * 2, st.cuda_stream)
    with torch.cuda.stream(st):
        for _ in range(2): run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(3): run()
        e1.record(st)
    e1.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"{W}x{H} nd {nd} mode {mode} batch {B}: {ms:.2f} ms/batch, {ms / B * 1e3:.0f} us/frame, {W * H * nd * B / ms / 1e3:.0f} Mde/s, launches {m.last_launches()}")
    del m
