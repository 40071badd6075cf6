"""Quick device-resident timing of the BM pipeline (development aid; bench.py is the contract)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rt-depth-map_b200"))
import numpy as np, torch
import rtdm_b200 as rt
from rtdm_b200 import synth

W, H, nd, B = 1280, 720, 128, int(sys.argv[1]) if len(sys.argv) > 1 else 32
print(rt.measure_int_peak(0))
frames = [synth.stereo_pair(W, H, nd, 1000 + i) for i in range(4)]
L = torch.from_numpy(np.stack([frames[i % 4][0] for i in range(B)])).cuda()
R = torch.from_numpy(np.stack([frames[i % 4][1] for i in range(B)])).cuda()
D = torch.empty((B, H, W), dtype=torch.int16, device="cuda")
m = rt.CUDAMatcherKonolige(None, None, 31, 13, 0, 10, nd, nd, 10, 100, 32, 1, max_width=W, max_height=H, max_batch=B)
st = torch.cuda.Stream()
def run():
    m.compute_device(B, L.data_ptr(), W, W * H, R.data_ptr(), W, W * H, W, H, D.data_ptr(), W * 2, W * H * 2, st.cuda_stream)
with torch.cuda.stream(st):
    for _ in range(3): run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(5): run()
    e1.record(st)
e1.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"batch {B}: {ms:.3f} ms/batch, {ms / B * 1e3:.1f} us/frame, {B / ms * 1e3:.0f} fps, {W * H * nd * B / ms / 1e3:.0f} Mde/s, launches {m.last_launches()}")
