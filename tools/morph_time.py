"""Development aid: device time of the open/close filter alone."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rt-depth-map_b200"))
import numpy as np, torch
import rtdm_b200 as rt
from rtdm_b200 import synth
W, H = 1280, 720
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
M = torch.from_numpy(np.stack([synth.binary_mask(W, H, 3000 + i % 4) for i in range(B)])).cuda()
O = torch.empty_like(M)
f = rt.CUDAMorphologicalFilter(W, H, 8, max_batch=B)
st = torch.cuda.Stream()
def run(): f.run_device(B, M.data_ptr(), O.data_ptr(), st.cuda_stream)
with torch.cuda.stream(st):
    for _ in range(3): run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(10): run()
    e1.record(st)
e1.synchronize()
print(f"morph batch {B}: {e0.elapsed_time(e1) / 10 / B * 1e3:.2f} us/frame, launches {f.last_launches()}")
