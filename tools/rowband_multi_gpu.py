"""One 1280x720 (or WxH) frame split into row bands over the GPUs of one box (torchrun, NCCL gather to rank 0):
checks the result against the whole-frame matcher on rank 0 and times both.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/rowband_multi_gpu.py [W H nd]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rt-depth-map_b200"))
import numpy as np, torch, torch.distributed as dist
import rtdm_b200 as rt
from rtdm_b200 import rowband, synth

W, H, nd = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (1280, 720, 128)
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
L, R, _ = synth.stereo_pair(W, H, nd, 1000)
dL, dR = torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda()
rb = rowband.RowBandKonolige(rt, W, H, 31, 13, 10, nd, 10, 100, 32, 1, dist=dist, device=local)
out = rb.compute(dL, dR)                      # warm-up (NCCL communicator, kernels)
for _ in range(3):
    rb.compute(dL, dR)
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
N = 20
for _ in range(N):
    out = rb.compute(dL, dR)
torch.cuda.synchronize(); dist.barrier()
ms_band = (time.perf_counter() - t0) / N * 1e3
if rank == 0:
    whole = rt.CUDAMatcherKonolige(None, None, 31, 13, 0, 10, nd, nd, 10, 100, 32, 1, max_width=W, max_height=H, device=local)
    D = torch.empty((H, W), dtype=torch.int16, device="cuda")
    st = torch.cuda.Stream()
    def run():
        whole.compute_device(1, dL.data_ptr(), W, W * H, dR.data_ptr(), W, W * H, W, H, D.data_ptr(), W * 2, W * H * 2, st.cuda_stream)
        st.synchronize()
    for _ in range(3): run()
    t0 = time.perf_counter()
    for _ in range(N): run()
    ms_whole = (time.perf_counter() - t0) / N * 1e3
    same = bool(torch.equal(out, D))
    print(f"rowband {W}x{H} nd={nd} world={world}: bit-exact={same} band-split {ms_band:.3f} ms/frame, single GPU {ms_whole:.3f} ms/frame")
dist.destroy_process_group()
