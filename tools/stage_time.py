"""Development aid: per-stage device time of the BM pipeline (ms per batch) under the current env."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rt-depth-map_b200"))
import numpy as np, torch
import rtdm_b200 as rt
from rtdm_b200 import synth
W, H, nd = 1280, 720, 128
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
frames = [synth.stereo_pair(W, H, nd, 1000 + i) for i in range(4)]
L = torch.from_numpy(np.stack([frames[i % 4][0] for i in range(B)])).cuda()
R = torch.from_numpy(np.stack([frames[i % 4][1] for i in range(B)])).cuda()
D = torch.empty((B, H, W), dtype=torch.int16, device="cuda")
m = rt.CUDAMatcherKonolige(None, None, 31, 13, 0, 10, nd, nd, 10, 100, 32, 1, max_width=W, max_height=H, max_batch=B)
st = torch.cuda.Stream()
def run():
    m.compute_device(B, L.data_ptr(), W, W * H, R.data_ptr(), W, W * H, W, H, D.data_ptr(), W * 2, W * H * 2, st.cuda_stream)
for _ in range(3): run()
torch.cuda.synchronize()
m.set_profiling(True)
for _ in range(5): run()
t, c = m.stage_times()
print(os.environ.get("RTDM_BM_DEBUG", "-"), {k: round(v / c / B * 1e3, 1) for k, v in t.items()}, "us/frame; kernel", m.last_kernel())
