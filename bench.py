#!/usr/bin/env python
"""bench.py -- throughput of the stereo hot path (BlockMatcher back-end + morphological filter).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload bm720|sgbm720]

One "step" = one pass of the hot path over one batch of synthetic rectified frames:
  bm720   (default; BASELINE.json configs[2], the configuration the metric is quoted on):
          Konolige BM 1280x720, numDisparities 128, the reference's parameters (main.cpp:134-135:
          cap 31, bs 13, tex 10, uniq 10, speckle 100/32, disp12 1) = prefilter + SAD/WTA + validate +
          mask + speckle, plus SWMorphologicalFilter's open+close on one 1280x720 mask per frame.
  sgbm720 (configs[3]): SGBM 1280x720 nd 128, 8-path MODE_HH.
Metric: Mde/s = W*H*numDisparities*frames / s / 1e6; whole-job aggregate.

`value`  : device-resident inputs/outputs, CUDA events on the launching stream, max over ranks.
`e2e`    : the same batch through the host-pointer C ABI call (rtdm_*_compute_batch / morph_run) with
           pinned HOST buffers; H2D and D2H copies inside the timed region.
Multi-GPU: frames are independent -> each rank processes its own batch (weak scaling, no collective
           on the data path); only the timing reduction uses torch.distributed.
`--impl reference`: the reference's own CPU implementation of the path (cv2 = the OpenCV routines that
           bm-sw.cpp / sgbm-sw.cpp / mf-sw.cpp call; falls back to the C oracle port if cv2 is
           missing) on the host cores, bounded sample per step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rt-depth-map_b200"))

import numpy as np  # noqa: E402

W, H, ND = 1280, 720, 128
BM_PARAMS = dict(preFilterCap=31, blockSize=13, minDisparity=0, textureThreshold=10, numDisparities=ND,
                 uniquenessRatio=10, speckleWindowSize=100, speckleRange=32, disp12MaxDiff=1)
SGBM_PARAMS = dict(blockSize=5, minDisparity=0, numDisparities=ND, uniquenessRatio=10,
                   speckleWindowSize=100, speckleRange=32, disp12MaxDiff=1, mode=1)
N_DISTINCT = 8          # distinct synthetic frames (seeds 1000..); the batch cycles through them
OPS_PER_DE = {"bm720": 8, "sgbm720": 96}       # SURVEY.md 8(d) algorithmic integer ops per de
HBM_BYTES_PER_FRAME = {"bm720": 4 * W * H, "sgbm720": 4 * (W - ND) * H * ND + 4 * W * H}


def mde_per_frame():
    return W * H * ND / 1e6


def make_frames(n):
    from rtdm_b200 import synth
    fr = [synth.stereo_pair(W, H, ND, 1000 + i) for i in range(min(n, N_DISTINCT))]
    masks = [synth.binary_mask(W, H, 3000 + i) for i in range(min(n, N_DISTINCT))]
    L = np.stack([fr[i % len(fr)][0] for i in range(n)])
    R = np.stack([fr[i % len(fr)][1] for i in range(n)])
    M = np.stack([masks[i % len(masks)] for i in range(n)])
    return L, R, M


# ---------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md: the clocks line)
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # median of the samples under load (upper half of the observed clocks)
        s = sorted(sm)
        under = s[len(s) // 2:]
        return {"sm_mhz": float(np.median(under)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation of the path
# ---------------------------------------------------------------------------------------------------
def cpu_runner(workload):
    """-> (fn(L, R, M) processing ONE frame, kind, cores, description)"""
    from oracle import cv2_ref
    cores = os.cpu_count() or 1
    if cv2_ref.have_cv2():
        import cv2
        cv2.setNumThreads(cores)
        if workload == "bm720":
            m = cv2_ref.make_bm(**BM_PARAMS)

            def fn(L, R, M):
                d = m.compute(L, R)
                o = cv2_ref.morph_open_close(M)
                return d, o
        else:
            m = cv2_ref.make_sgbm(**SGBM_PARAMS)

            def fn(L, R, M):
                return m.compute(L, R), None
        return fn, "reference", cores, f"cv2 {cv2.__version__} (OpenCV routines the reference calls), {cores} threads"
    from oracle import oracle
    if workload == "bm720":
        p = oracle.make_params(**BM_PARAMS)

        def fn(L, R, M):
            return oracle.bm_compute(L, R, p), oracle.morph_open_close(M)
    else:
        p = oracle.make_params(P1=600, P2=2400, preFilterCap=0, **SGBM_PARAMS)

        def fn(L, R, M):
            return oracle.sgbm_compute(L, R, p), None
    return fn, "port", 1, "oracle/stereo_oracle.c (scalar C restatement), 1 thread"


def time_cpu(workload, frames_per_step, steps, warmup):
    fn, kind, cores, desc = cpu_runner(workload)
    L, R, M = make_frames(min(frames_per_step, N_DISTINCT))
    n = L.shape[0]
    for i in range(max(1, warmup)):
        fn(L[i % n], R[i % n], M[i % n])
    t0 = time.perf_counter()
    for s in range(steps):
        for i in range(frames_per_step):
            fn(L[i % n], R[i % n], M[i % n])
    dt = time.perf_counter() - t0
    return dict(seconds=dt, frames=steps * frames_per_step, kind=kind, cores=cores, desc=desc)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fps_guess = {"bm720": 25.0, "sgbm720": 1.2}[args.workload]
    # bounded sample: keep the whole run within ~a minute
    frames_per_step = max(1, int(round(min(8, 30.0 * fps_guess / max(1, args.steps + args.warmup)))))
    r = time_cpu(args.workload, frames_per_step, args.steps, args.warmup)
    fps = r["frames"] / r["seconds"]
    value = fps * mde_per_frame()
    line = {
        "impl": "reference", "metric": "Mde/s", "value": value, "unit": "Mde/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["seconds"] / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8/u16 integer", "data": "synthetic",
        "fps": fps,
        "config": {"workload": workload_name(args.workload), "frames_per_step": frames_per_step,
                   "width": W, "height": H, "numDisparities": ND},
        "cpu_baseline": {"value": value, "unit": "Mde/s", "cores": r["cores"], "kind": r["kind"],
                         "sample": f"{r['frames']} frames of the workload, {r['desc']}"},
        "e2e": {"value": value, "unit": "Mde/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_name(w):
    return {"bm720": "Konolige BM 1280x720 nd=128 bs=13 (prefilter+SAD/WTA+uniqueness+disp12+speckle) + 10x10 ellipse open/close",
            "sgbm720": "SGBM 1280x720 nd=128 bs=5 P1=600 P2=2400 MODE_HH 8-path + median + speckle"}[w]


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import rtdm_b200 as rt

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or rt.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    wl = args.workload
    # SGBM keeps 2 x 212 MB of cost volumes per frame; 37 frames x 24 column tiles = 888 sweep CTAs = 6.0 waves of 148 SMs
    B = args.batch if wl == "bm720" else min(args.batch, 37)
    Lh, Rh, Mh = make_frames(B)
    dev = torch.device("cuda", local)
    L, R, M = (torch.from_numpy(a).to(dev) for a in (Lh, Rh, Mh))
    D = torch.empty((B, H, W), dtype=torch.int16, device=dev)
    MO = torch.empty_like(M)
    if wl == "bm720":
        matcher = rt.CUDAMatcherKonolige(None, None, BM_PARAMS["preFilterCap"], BM_PARAMS["blockSize"], 0,
                                         BM_PARAMS["textureThreshold"], ND, ND, BM_PARAMS["uniquenessRatio"],
                                         BM_PARAMS["speckleWindowSize"], BM_PARAMS["speckleRange"],
                                         BM_PARAMS["disp12MaxDiff"], max_width=W, max_height=H, max_batch=B, device=local)
        filt = rt.CUDAMorphologicalFilter(W, H, 8, max_batch=B, device=local)
    else:
        matcher = rt.CUDASemiGlobalMatcher(SGBM_PARAMS["blockSize"], 0, ND, SGBM_PARAMS["uniquenessRatio"],
                                           SGBM_PARAMS["speckleWindowSize"], SGBM_PARAMS["speckleRange"],
                                           SGBM_PARAMS["disp12MaxDiff"], mode=SGBM_PARAMS["mode"],
                                           max_width=W, max_height=H, max_batch=B, device=local)
        filt = None
    st = torch.cuda.Stream(device=dev)

    def step():
        matcher.compute_device(B, L.data_ptr(), W, W * H, R.data_ptr(), W, W * H, W, H,
                               D.data_ptr(), W * 2, W * H * 2, st.cuda_stream)
        if filt is not None:
            filt.run_device(B, M.data_ptr(), MO.data_ptr(), st.cuda_stream)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()         # nvidia-smi needs ~100 ms to deliver its first sample: start before warm-up
    with torch.cuda.stream(st):
        for _ in range(args.warmup):
            step()
    launches_per_step = matcher.last_launches() + (filt.last_launches() if filt is not None else 0)
    if hasattr(matcher, "set_profiling"):
        matcher.set_profiling(True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record(st)
        for _ in range(args.steps):
            step()
        e1.record(st)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    stage_ms, stage_calls = ({}, 0)
    if hasattr(matcher, "stage_times"):
        stage_ms, stage_calls = matcher.stage_times()
        matcher.set_profiling(False)
    from rtdm_b200 import sharding
    # frames are independent: every rank processed its own B*steps frames; value = all frames / max time
    fps, ms_max, frames_total = sharding.whole_job_throughput(B * args.steps, ms, 1.0, dist, dev)
    value = fps * mde_per_frame()

    # ---- end to end through the host-pointer C ABI with pinned host buffers ------------------------
    e2e_steps = max(1, args.steps)          # the same K steps; the stream's fill and drain are inside the timed region
    Lp, Rp, Mp = (torch.from_numpy(a).pin_memory() for a in (Lh, Rh, Mh))
    Dp = torch.empty((B, H, W), dtype=torch.int16).pin_memory()
    Lpn, Rpn, Dpn = Lp.numpy(), Rp.numpy(), Dp.numpy()
    Mpn = Mp.numpy()
    MOpn = torch.empty((B, H, W), dtype=torch.uint8).pin_memory().numpy()

    streaming = hasattr(matcher, "submit_batch")
    Dpn2 = torch.empty((B, H, W), dtype=torch.int16).pin_memory().numpy() if streaming else None
    acc = [0]

    filter_first = bool(os.environ.get("RTDM_BENCH_FILTER_FIRST"))

    def e2e_step(i=0):
        # the filter's copies and kernels run on its own stream underneath the matcher call; the matcher's copies
        # are enqueued first because its kernels (the long pole) cannot start before their first chunk has arrived.
        # Streaming use of both plugins: every batch's results (disparity maps AND filtered masks) are read on the host
        # one submission later, so the host never blocks on work it has only just enqueued.
        if filt is not None and filter_first:
            if i > 0:
                filt.sync(); acc[0] += int(MOpn[0, H // 2, W // 2])
            filt.run_batch_async(Mpn, MOpn)
        if streaming:
            # depth-2 stream of batches: the copies of batch i+1 / i-1 run under the kernels of batch i
            matcher.submit_batch(Lpn, Rpn, Dpn2 if i & 1 else Dpn)
            if filt is not None and not filter_first:
                if i > 0:
                    filt.sync(); acc[0] += int(MOpn[0, H // 2, W // 2])
                filt.run_batch_async(Mpn, MOpn)
            if i > 0:
                matcher.wait_oldest()
                acc[0] += int((Dpn if i & 1 else Dpn2)[0, H // 2, W // 2])
        else:
            if filt is not None and not filter_first:
                filt.run_batch_async(Mpn, MOpn)
            matcher.compute_batch(Lpn, Rpn, Dpn)
            acc[0] += int(Dpn[0, H // 2, W // 2])
            if filt is not None:
                filt.sync()
                acc[0] += int(MOpn[0, H // 2, W // 2])

    e2e_step()
    if streaming:
        matcher.wait()
        if filt is not None:
            filt.sync()
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(i)
    if streaming:
        matcher.wait()          # the last batch's results land inside the timed region
        acc[0] += int((Dpn2 if (e2e_steps - 1) & 1 else Dpn)[0, H // 2, W // 2])
        if filt is not None:
            filt.sync()
            acc[0] += int(MOpn[0, H // 2, W // 2])
    torch.cuda.synchronize()
    e2e_fps, _, _ = sharding.whole_job_throughput(B * e2e_steps, (time.perf_counter() - t0) * 1e3, 1.0, dist, dev)
    h2d = B * 2 * W * H + (B * W * H if filt is not None else 0)
    d2h = B * W * H * 2 + (B * W * H if filt is not None else 0)

    if rank == 0:
        # ---- roofline of the dominant kernel ------------------------------------------------------
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        roofline = None
        if stage_calls and wl == "bm720":
            k_ms = stage_ms["sad_wta"] / stage_calls            # one launch = B frames
            ip = rt.measure_int_peak(local)
            int_peak = ip["iadd3_tiops"]
            de = B * W * H * ND
            ach = de * OPS_PER_DE[wl] / (k_ms * 1e-3) / 1e12
            hbm_ach = B * HBM_BYTES_PER_FRAME[wl] / (k_ms * 1e-3) / 1e9
            roofline = {
                "kernel": {3: "bm_texture_kernel + bm_sad3_kernel (warp-specialised SAD/WTA)", 2: "bm_texture_kernel + bm_sad2_kernel",
                           1: "bm_sad_wta_kernel (generic)"}.get(matcher.last_kernel(), "?"),
                "bound": "int_alu", "achieved": ach, "peak": int_peak, "unit": "Tiop/s",
                "frac": ach / int_peak,
                # ncu dram__bytes_read.sum + dram__bytes_write.sum of one bm_sad3_kernel launch of 63 frames (capture H,
                # profiles/r01_prof_bm3_h_summary.csv): 235.5 + 172.9 MB = 6.48 MB per frame -- the kernel reads the two
                # prefiltered images and the texture sums (3.7 MB) and writes raw disparity + cost (3.7 MB, part of it still
                # in L2 when the launch ends); the 4 B/pixel of SURVEY 8(d) count the path's own inputs and output only
                "traffic": 6.48e6 * B if matcher.last_kernel() == 3 else (4.7e6 * B if matcher.last_kernel() == 2 else None),
                "traffic_note": "per launch of B frames, scaled from the 63-frame ncu capture under profiles/ (capture H)",
                "ops_per_de": OPS_PER_DE[wl], "kernel_ms_per_launch": k_ms, "frames_per_launch": B,
                "kernel_share_of_step": stage_ms["sad_wta"] / stage_calls / (ms / args.steps),
                "peak_source": "rtdm_measure_int_peak on this GPU (dependent-free IADD3, lane-ops/s)",
                "int_peak_detail": ip,
                "hbm": {"achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak,
                        "peak_source": hbm_src, "algorithmic_bytes_per_frame": HBM_BYTES_PER_FRAME[wl]},
                "stage_ms_per_step": {k: v / stage_calls for k, v in stage_ms.items()},
            }
        if stage_calls and wl == "sgbm720":
            # the matching stage as a whole (planes, fused cost, first path, row sweeps, last path + WTA, LR check):
            # SURVEY.md 8(d) rates it against the integer pipes (96 ops/de for MODE_HH); the HBM figures sit beside it
            k_ms = stage_ms["matching"] / stage_calls
            ip = rt.measure_int_peak(local)
            hbm_ach = B * HBM_BYTES_PER_FRAME[wl] / (k_ms * 1e-3) / 1e9
            ach_int = B * W * H * ND * OPS_PER_DE[wl] / (k_ms * 1e-3) / 1e12
            traffic = 2190e6 * B            # profiles/r01_launches_E_sgbm_hh.csv: 17.5 GB of DRAM traffic per 8 frames
            roofline = {
                "kernel": "sgbm matching stage (sgbm_sweep_kernel x2 passes dominant, + cost_fused, path4 first/last, lr)",
                "bound": "int_alu", "achieved": ach_int, "peak": ip["iadd3_tiops"], "unit": "Tiop/s", "frac": ach_int / ip["iadd3_tiops"],
                "traffic": traffic,
                "traffic_note": "ncu dram bytes summed over the stage's launches, profiles/r01_launches_E_sgbm_hh.csv: 2.19 GB per frame",
                "ops_per_de": OPS_PER_DE[wl], "stage_ms_per_launch": k_ms, "frames_per_launch": B,
                "peak_source": "rtdm_measure_int_peak on this GPU (dependent-free IADD3, lane-ops/s)",
                "int_peak_detail": ip,
                "hbm": {"achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak, "peak_source": hbm_src,
                        "algorithmic_bytes_per_frame": HBM_BYTES_PER_FRAME[wl],
                        "moved_gbs": traffic / (k_ms * 1e-3) / 1e9, "moved_frac_of_peak": traffic / (k_ms * 1e-3) / 1e9 / hbm_peak},
                "stage_ms_per_step": {k: v / stage_calls for k, v in stage_ms.items()},
            }
        # ---- CPU baseline on a bounded sample (rank 0, N=1 only) -----------------------------------
        cpu = None
        if world == 1 and not args.no_cpu:
            nfr = {"bm720": 240, "sgbm720": 12}[wl]
            r = time_cpu(wl, nfr, 1, 1)
            cfps = r["frames"] / r["seconds"]
            cpu = {"value": cfps * mde_per_frame(), "unit": "Mde/s", "cores": r["cores"], "kind": r["kind"],
                   "sample": f"{r['frames']} frames of the workload in {r['seconds']:.1f} s, {r['desc']}", "fps": cfps}
        line = {
            "metric": "Mde/s", "value": value, "unit": "Mde/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/u16 integer", "data": "synthetic", "fps": fps,
            "config": {"workload": workload_name(wl), "frames_per_step_per_gpu": B, "width": W, "height": H,
                       "numDisparities": ND, "parallelism": f"frame-sharded x{world} (no collective)",
                       "l2": f"inputs+outputs per step = {(B * 4 * W * H + 2 * B * W * H) / 1e6:.0f} MB > 126 MB L2"},
            "e2e": {"value": e2e_fps * mde_per_frame(), "unit": "Mde/s", "fps": e2e_fps, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "api": ("rtdm_bm_submit_batch (2 batches in flight) + rtdm_morph_run_batch_async, results of batch i read during batch i+1 (rtdm_bm_wait_oldest / rtdm_morph_sync)" if wl == "bm720" else "rtdm_sgbm_submit_batch (2 batches in flight), results of batch i read during batch i+1 (rtdm_sgbm_wait_oldest)") + ", pinned host buffers"},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="bm720", choices=["bm720", "sgbm720"])
    # 63: the SAD/WTA kernel runs 42 CTAs per 720p frame, one per SM -> 63 frames = 17.9 waves of 148 SMs (64: 18.2 -> 19)
    ap.add_argument("--batch", type=int, default=63, help="frames per step per GPU")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
