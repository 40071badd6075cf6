#!/usr/bin/env python
"""bench.py -- throughput of the stereo hot path (BlockMatcher back-end + morphological filter).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload all|bm720|sgbm720]

One "step" = one pass of the hot path over one batch of synthetic rectified frames (all frames of a batch distinct):
  bm720   (top level of the JSON line; BASELINE.json configs[2], the configuration the metric is quoted on):
          Konolige BM 1280x720, numDisparities 128, the reference's parameters (main.cpp:134-135:
          cap 31, bs 13, tex 10, uniq 10, speckle 100/32, disp12 1) = prefilter + SAD/WTA + validate +
          mask + speckle, plus SWMorphologicalFilter's open+close on one 1280x720 mask per frame.
  sgbm720 (the "sgbm" object of the same line; configs[3]): SGBM 1280x720 nd 128 bs 5 P1 600 P2 2400, once as
          MODE_HH (8 paths) and once as MODE_SGBM (5 paths, the reference's literal default, sgbm-sw.cpp:15).
  latency (the "latency" object): Estimator::run's real call pattern (estimator.cpp:45,54-56; main.cpp:131-135):
          one frame, synchronous, host pointers: filter run on a 934x404 mask, setROI1, BM compute on the 934x404
          ROI crop (strided views of 1280x720 images) with numDisparities 192 -- beside cv2 doing the same; plus
          the reference's other matcher on the same views (SWSemiGlobalMatcher::compute, MODE_SGBM, `sgbm_ms`).
Metric: Mde/s = W*H*numDisparities*frames / s / 1e6; whole-job aggregate.

`value`  : device-resident inputs/outputs, CUDA events on the launching stream, max over ranks; the K steps are
           repeated (`repeats`) until the timed region lasts >= 1 s, `ms_per_step` is the mean.
`e2e`    : the same batches through the host-pointer C ABI (rtdm_*_submit_batch / rtdm_morph_run_batch_async) with
           pinned HOST buffers; H2D and D2H copies inside the timed region (>= 1 s as well).
`parity_checked`: after the timed regions, frames of the e2e OUTPUT buffers are compared with cv2 (the OpenCV
           routines the reference calls; the C oracle if cv2 is missing) -- any mismatch fails the run.
Multi-GPU: frames are independent -> each rank processes its own batches (weak scaling, no collective
           on the data path); only the timing reduction uses torch.distributed.
`--impl reference`: the reference's own CPU implementation of the path (cv2 = the OpenCV routines that
           bm-sw.cpp / sgbm-sw.cpp / mf-sw.cpp call; falls back to the C oracle port if cv2 is
           missing) on the host cores, bounded sample per step.
"""
from __future__ import annotations

import argparse
import json
import math
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
                   speckleWindowSize=100, speckleRange=32, disp12MaxDiff=1)
SGBM_MODES = {"mode_hh": 1, "mode_sgbm": 0}
OPS_PER_DE = {"bm720": 8, "mode_hh": 96, "mode_sgbm": 72}       # SURVEY.md 8(d) algorithmic integer ops per de
HBM_BYTES_PER_FRAME = {"bm720": 4 * W * H, "mode_hh": 4 * (W - ND) * H * ND + 4 * W * H, "mode_sgbm": 4 * (W - ND) * H * ND + 4 * W * H}
MIN_REGION_S = 1.0          # every timed region lasts at least this long
BM_BATCH, SGBM_BATCH = 63, 60
# Estimator's real operating point (latency leg)
OPW, OPH, OPND, OPX, OPY = 934, 404, 192, 173, 158      # backup/1280x720/extrinsics.yml:56-57 via main.cpp:80-85; -nd 192

WORKLOAD_BM = "Konolige BM 1280x720 nd=128 bs=13 (prefilter+SAD/WTA+uniqueness+disp12+speckle) + 10x10 ellipse open/close"
WORKLOAD_SGBM = {"mode_hh": "SGBM 1280x720 nd=128 bs=5 P1=600 P2=2400 MODE_HH 8-path + median + speckle",
                 "mode_sgbm": "SGBM 1280x720 nd=128 bs=5 P1=600 P2=2400 MODE_SGBM 5-path + median + speckle"}


def mde_per_frame():
    return W * H * ND / 1e6


def config_for(world):
    """Identical in both arms (`--impl ours` / `--impl reference`): the driver compares the two dicts."""
    return {"workload": WORKLOAD_BM, "frames_per_step_per_gpu": BM_BATCH, "width": W, "height": H, "numDisparities": ND,
            "parallelism": f"frame-sharded x{world} (no collective)",
            "distinct_frames_per_gpu": BM_BATCH,
            "l2": f"inputs+outputs per step = {(BM_BATCH * 4 * W * H + 2 * BM_BATCH * W * H) / 1e6:.0f} MB > 126 MB L2",
            "sgbm": {"workloads": WORKLOAD_SGBM, "frames_per_step_per_gpu": SGBM_BATCH},
            "latency": f"one frame, synchronous, host pointers: open/close on a {OPW}x{OPH} mask, setROI1, BM {OPW}x{OPH} nd={OPND} bs=13 on strided ROI views"}


def make_frames(n, seed0=1000, masks=True):
    """n DISTINCT synthetic rectified pairs (frame i uses seed 1000 + i) and binary masks."""
    from rtdm_b200 import synth
    L = np.empty((n, H, W), np.uint8); R = np.empty((n, H, W), np.uint8)
    M = np.empty((n, H, W), np.uint8) if masks else None
    for i in range(n):
        L[i], R[i], _ = synth.stereo_pair(W, H, ND, seed0 + i)
        if masks:
            M[i] = synth.binary_mask(W, H, seed0 + 2000 + i)
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
        return self

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
        sm, mx, pw, reasons = [], [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            try:
                pw.append(float(f[3]))
            except ValueError:
                pass
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # median of the samples under load (upper half of the observed clocks)
        s = sorted(sm)
        under = s[len(s) // 2:]
        return {"sm_mhz": float(np.median(under)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(pw) if pw else None}


# ---------------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation of the path
# ---------------------------------------------------------------------------------------------------
def cpu_runner(workload, threads=None):
    """-> (fn(L, R, M) processing ONE frame, kind, cores, description).  workload: bm720 | mode_hh | mode_sgbm"""
    from oracle import cv2_ref
    cores = threads or os.cpu_count() or 1
    if cv2_ref.have_cv2():
        import cv2
        cv2.setNumThreads(cores)
        if workload == "bm720":
            m = cv2_ref.make_bm(**BM_PARAMS)

            def fn(L, R, M):
                d = m.compute(L, R)
                o = cv2_ref.morph_open_close(M) if M is not None else None
                return d, o
        else:
            m = cv2_ref.make_sgbm(mode=SGBM_MODES[workload], **SGBM_PARAMS)

            def fn(L, R, M):
                return m.compute(L, R), None
        return fn, "reference", cores, f"cv2 {cv2.__version__} (OpenCV routines the reference calls), {cores} thread{'s' if cores > 1 else ''}"
    from oracle import oracle
    if workload == "bm720":
        p = oracle.make_params(**BM_PARAMS)

        def fn(L, R, M):
            return oracle.bm_compute(L, R, p), (oracle.morph_open_close(M) if M is not None else None)
    else:
        p = oracle.make_params(P1=600, P2=2400, preFilterCap=0, mode=SGBM_MODES[workload], **SGBM_PARAMS)

        def fn(L, R, M):
            return oracle.sgbm_compute(L, R, p), None
    return fn, "port", 1, "oracle/stereo_oracle.c (scalar C restatement), 1 thread"


def time_cpu(workload, frames, total, warmup=1, threads=None):
    """Times `total` frames (cycling through the distinct `frames`) after `warmup` untimed ones."""
    fn, kind, cores, desc = cpu_runner(workload, threads)
    L, R, M = frames
    n = L.shape[0]
    for i in range(max(1, warmup)):
        fn(L[i % n], R[i % n], M[i % n] if M is not None else None)
    t0 = time.perf_counter()
    for i in range(total):
        fn(L[i % n], R[i % n], M[i % n] if M is not None else None)
    dt = time.perf_counter() - t0
    return dict(seconds=dt, frames=total, kind=kind, cores=cores, desc=desc, fps=total / dt)


def cpu_baseline_obj(workload, frames, n_all, n_one):
    """cpu_baseline object: all host threads, plus the 1-thread figure BASELINE.md section 3 asks for."""
    r = time_cpu(workload, frames, n_all, 1)
    out = {"value": r["fps"] * mde_per_frame(), "unit": "Mde/s", "cores": r["cores"], "kind": r["kind"], "fps": r["fps"],
           "sample": f"{r['frames']} frames of the workload in {r['seconds']:.1f} s, {r['desc']}"}
    if r["kind"] == "reference" and n_one > 0:
        r1 = time_cpu(workload, frames, n_one, 1, threads=1)
        out["one_thread"] = {"value": r1["fps"] * mde_per_frame(), "fps": r1["fps"], "sample": f"{r1['frames']} frames in {r1['seconds']:.1f} s"}
        import cv2
        cv2.setNumThreads(os.cpu_count() or 1)
    return out


def op_rois():
    """setROI1 rectangles of the latency leg (what find_relevant_matching_region would hand over, estimator.cpp:53-54)."""
    return [(260, 60, 520, 280), (0, 0, OPW, OPH), (400, 100, 300, 200), (600, 10, 334, 390), (10, 300, 900, 104),
            (200, 40, 640, 330), (300, 150, 400, 200), (120, 20, 760, 360)]


def op_frames(n=4):
    """Full 1280x720 images whose calibrated ROI crop holds a synthetic rectified pair with up to OPND disparities."""
    from rtdm_b200 import synth
    Ls, Rs, Ms = [], [], []
    for i in range(n):
        l, r, _ = synth.stereo_pair(OPW, OPH, OPND, 5000 + i)
        fl = np.zeros((H, W), np.uint8); fr = np.zeros((H, W), np.uint8)
        fl[OPY:OPY + OPH, OPX:OPX + OPW] = l; fr[OPY:OPY + OPH, OPX:OPX + OPW] = r
        Ls.append(fl); Rs.append(fr); Ms.append(synth.binary_mask(OPW, OPH, 5100 + i))
    return Ls, Rs, Ms


def cpu_latency(iters=12):
    """cv2 doing Estimator's per-frame calls: filter run, setROI1, compute on the strided ROI views."""
    from oracle import cv2_ref
    Ls, Rs, Ms = op_frames()
    rois = op_rois()
    if not cv2_ref.have_cv2():
        return None
    import cv2
    cv2.setNumThreads(os.cpu_count() or 1)
    m = cv2_ref.make_bm(**dict(BM_PARAMS, numDisparities=OPND))
    tb = tf = 0.0
    for i in range(-2, iters):
        k = i % len(Ls)
        Lv, Rv = Ls[k][OPY:OPY + OPH, OPX:OPX + OPW], Rs[k][OPY:OPY + OPH, OPX:OPX + OPW]
        t0 = time.perf_counter()
        cv2_ref.morph_open_close(Ms[k])
        t1 = time.perf_counter()
        m.setROI1(rois[i % len(rois)])
        m.compute(Lv, Rv)
        t2 = time.perf_counter()
        if i >= 0:
            tf += t1 - t0; tb += t2 - t1
    sg = cv2_ref.make_sgbm(mode=SGBM_MODES["mode_sgbm"], **dict(SGBM_PARAMS, numDisparities=OPND))
    ts, sg_iters = 0.0, 3
    for i in range(-1, sg_iters):
        k = i % len(Ls)
        t0 = time.perf_counter()
        sg.compute(Ls[k][OPY:OPY + OPH, OPX:OPX + OPW], Rs[k][OPY:OPY + OPH, OPX:OPX + OPW])
        if i >= 0:
            ts += time.perf_counter() - t0
    return {"filter_ms": tf / iters * 1e3, "bm_ms": tb / iters * 1e3, "frame_ms": (tf + tb) / iters * 1e3, "frames": iters,
            "sgbm_ms": ts / sg_iters * 1e3, "sgbm_frames": sg_iters,
            "impl": f"cv2 {cv2.__version__}, {os.cpu_count()} threads"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    K, Wm = max(1, args.steps), max(1, args.warmup)
    # bounded sample per step: keep the whole run within ~a minute
    fps_guess = 25.0
    frames_per_step = max(1, int(round(min(8, 30.0 * fps_guess / (K + Wm)))))
    frames = make_frames(min(16, frames_per_step * 2))
    r = time_cpu("bm720", frames, K * frames_per_step, Wm * frames_per_step)
    value = r["fps"] * mde_per_frame()
    line = {
        "impl": "reference", "metric": "Mde/s", "value": value, "unit": "Mde/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["seconds"] / K * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8/u16 integer", "data": "synthetic", "fps": r["fps"],
        "config": config_for(args.gpus),
        "cpu_baseline": {"value": value, "unit": "Mde/s", "cores": r["cores"], "kind": r["kind"],
                         "sample": f"{frames_per_step} frames per step ({r['frames']} frames) of the workload, {r['desc']}"},
        "e2e": {"value": value, "unit": "Mde/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if args.workload in ("all", "sgbm720"):
        sg = {}
        fr2 = (frames[0][:4], frames[1][:4], None)
        for name in SGBM_MODES:
            rs = time_cpu(name, fr2, min(K, 8), 1)
            v = rs["fps"] * mde_per_frame()
            sg[name] = {"workload": WORKLOAD_SGBM[name], "value": v, "unit": "Mde/s", "fps": rs["fps"],
                        "ms_per_step": rs["seconds"] / rs["frames"] * 1e3,
                        "cpu_baseline": {"value": v, "unit": "Mde/s", "cores": rs["cores"], "kind": rs["kind"],
                                         "sample": f"1 frame per step, {rs['frames']} frames, {rs['desc']}"},
                        "e2e": {"value": v, "unit": "Mde/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        line["sgbm"] = sg
    if args.workload == "all":
        line["latency"] = {"reference": cpu_latency()}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
class Ctx:
    pass


def timed_region(ctx, step, st, K):
    """Runs K steps once to estimate their duration, then times K x repeats steps in ONE region of >= MIN_REGION_S
    (CUDA events on the launching stream, barrier + synchronize on both sides).  -> (ms of the region, repeats)"""
    torch = ctx.torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    with torch.cuda.stream(st):
        e0.record(st)
        for _ in range(K):
            step()
        e1.record(st)
    ctx.barrier()
    est = max(e0.elapsed_time(e1), 1e-3)
    from rtdm_b200 import sharding
    est = sharding.max_over_ranks(est, ctx.dist, ctx.dev)            # the same repeat count on every rank
    repeats = max(1, int(math.ceil(MIN_REGION_S * 1e3 * 1.05 / est)))
    ctx.before_timed()
    ctx.barrier()
    with torch.cuda.stream(st):
        e0.record(st)
        for _ in range(K * repeats):
            step()
        e1.record(st)
    ctx.barrier()
    return e0.elapsed_time(e1), repeats


def link_ceiling(ctx):
    """Host link measured the way the e2e legs use it: pinned 64 MB blocks, H2D and D2H at the same time on two
    streams, all ranks at once.  -> GB/s per direction on this rank."""
    torch = ctx.torch
    n = 64 << 20
    h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); d1 = torch.empty(n, dtype=torch.uint8, device=ctx.dev)
    h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.zeros(n, dtype=torch.uint8, device=ctx.dev)
    s1, s2 = torch.cuda.Stream(device=ctx.dev), torch.cuda.Stream(device=ctx.dev)

    def go(k):
        for _ in range(k):
            with torch.cuda.stream(s1):
                d1.copy_(h1, non_blocking=True)
            with torch.cuda.stream(s2):
                h2.copy_(d2, non_blocking=True)
    go(2)
    ctx.barrier()
    t0 = time.perf_counter()
    go(12)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ctx.barrier()
    return 12 * n / dt / 1e9


def load_traffic():
    """ncu dram bytes per frame of the dominant kernels, from the captures committed under profiles/ (traffic.json
    names the capture each figure comes from)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        return {}


def bench_bm(ctx, K, frames):
    torch, rt = ctx.torch, ctx.rt
    from rtdm_b200 import sharding
    B = BM_BATCH
    Lh, Rh, Mh = frames
    dev, local = ctx.dev, ctx.local
    L, R, M = (torch.from_numpy(a).to(dev) for a in (Lh, Rh, Mh))
    D = torch.empty((B, H, W), dtype=torch.int16, device=dev)
    MO = torch.empty_like(M)
    matcher = rt.CUDAMatcherKonolige(None, None, BM_PARAMS["preFilterCap"], BM_PARAMS["blockSize"], 0,
                                     BM_PARAMS["textureThreshold"], ND, ND, BM_PARAMS["uniquenessRatio"],
                                     BM_PARAMS["speckleWindowSize"], BM_PARAMS["speckleRange"],
                                     BM_PARAMS["disp12MaxDiff"], max_width=W, max_height=H, max_batch=B, device=local)
    filt = rt.CUDAMorphologicalFilter(W, H, 8, max_batch=B, device=local)
    st = torch.cuda.Stream(device=dev)

    def step():
        matcher.compute_device(B, L.data_ptr(), W, W * H, R.data_ptr(), W, W * H, W, H,
                               D.data_ptr(), W * 2, W * H * 2, st.cuda_stream)
        filt.run_device(B, M.data_ptr(), MO.data_ptr(), st.cuda_stream)

    sampler = ClockSampler(local).start() if ctx.rank == 0 else None     # nvidia-smi needs ~100 ms for its first sample
    with torch.cuda.stream(st):
        for _ in range(ctx.args.warmup):
            step()
    launches_per_step = matcher.last_launches() + filt.last_launches()
    ctx.before_timed = lambda: matcher.set_profiling(True)
    ms, repeats = timed_region(ctx, step, st, K)
    nsteps = K * repeats
    clocks = sampler.stop() if sampler else None
    stage_ms, stage_calls = matcher.stage_times()
    matcher.set_profiling(False)
    fps, ms_max, _ = sharding.whole_job_throughput(B * nsteps, ms, 1.0, ctx.dist, dev)
    # device-side result of the last step, for the parity check
    dev_out = (D[[0, B // 2, B - 1]].cpu().numpy(), MO[[0]].cpu().numpy())

    # ---- end to end through the host-pointer C ABI with pinned host buffers ------------------------
    Lp, Rp, Mp = (torch.from_numpy(a).pin_memory().numpy() for a in (Lh, Rh, Mh))
    Dp = [torch.empty((B, H, W), dtype=torch.int16).pin_memory().numpy() for _ in range(2)]
    MOp = torch.empty((B, H, W), dtype=torch.uint8).pin_memory().numpy()
    acc = [0]

    def e2e_step(i):
        # depth-2 stream of batches (the copies of batch i+1 / i-1 run under the kernels of batch i); the filter's copies
        # and kernels run on its own stream.  Every batch's results (disparity maps AND filtered masks) are read on the
        # host one submission later, so the host never blocks on work it has only just enqueued.
        matcher.submit_batch(Lp, Rp, Dp[i & 1])
        if i > 0:
            filt.sync(); acc[0] += int(MOp[0, H // 2, W // 2])
        filt.run_batch_async(Mp, MOp)
        if i > 0:
            matcher.wait_oldest()
            acc[0] += int(Dp[(i - 1) & 1][0, H // 2, W // 2])

    def e2e_run(n):
        for i in range(n):
            e2e_step(i)
        matcher.wait()              # the last batch's results land inside the timed region
        filt.sync()
        acc[0] += int(Dp[(n - 1) & 1][0, H // 2, W // 2]) + int(MOp[0, H // 2, W // 2])

    e2e_run(2)
    ctx.barrier()
    t0 = time.perf_counter()
    e2e_run(K)
    est = sharding.max_over_ranks(time.perf_counter() - t0, ctx.dist, dev)
    e2e_steps = K * max(1, int(math.ceil(MIN_REGION_S * 1.05 / max(est, 1e-6))))
    for d in Dp:
        d[:] = 0x5555
    MOp[:] = 0x55
    ctx.barrier()
    t0 = time.perf_counter()
    e2e_run(e2e_steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_fps, e2e_ms_max, _ = sharding.whole_job_throughput(B * e2e_steps, e2e_s * 1e3, 1.0, ctx.dist, dev)
    h2d = B * 3 * W * H
    d2h = B * W * H * 2 + B * W * H
    link = ctx.link_gbs
    gbs_dir = max(h2d, d2h) * e2e_steps / e2e_s / 1e9               # this rank, the busier direction

    res = {"value": fps * mde_per_frame(), "fps": fps, "ms_per_step": ms_max / nsteps, "repeats": repeats,
           "timed_region_s": ms_max * 1e-3, "clocks": clocks, "gpu_launches": launches_per_step * nsteps,
           "e2e": {"value": e2e_fps * mde_per_frame(), "unit": "Mde/s", "fps": e2e_fps, "h2d_bytes_per_step": h2d,
                   "d2h_bytes_per_step": d2h, "steps": e2e_steps, "timed_region_s": e2e_ms_max * 1e-3,
                   "link_ceiling_gbs": link, "link_gbs": gbs_dir, "link_frac": gbs_dir / link if link else None,
                   "link_note": "per rank and direction; ceiling = pinned 64 MB blocks, H2D and D2H at once, all ranks at once",
                   "api": "rtdm_bm_submit_batch (2 batches in flight) + rtdm_morph_run_batch_async, results of batch i read during "
                          "batch i+1 (rtdm_bm_wait_oldest / rtdm_morph_sync), pinned host buffers"},
           "_out": (Dp[(e2e_steps - 1) & 1], MOp, dev_out), "_stage": (stage_ms, stage_calls), "_kernel": matcher.last_kernel()}
    if ctx.rank == 0 and stage_calls:
        k_ms = stage_ms["sad_wta"] / stage_calls            # one launch = B frames
        ip = ctx.int_peak
        de = B * W * H * ND
        ach = de * OPS_PER_DE["bm720"] / (k_ms * 1e-3) / 1e12
        hbm_ach = B * HBM_BYTES_PER_FRAME["bm720"] / (k_ms * 1e-3) / 1e9
        tr = ctx.traffic.get({3: "bm_sad3", 4: "bm_sad4"}.get(matcher.last_kernel(), "-"), {})
        step_ms = ms / nsteps
        res["roofline"] = {
            "kernel": {4: "bm_sad4_kernel (TMA-staged, warp-specialised SAD/WTA)", 3: "bm_sad3_kernel (warp-specialised SAD/WTA)", 2: "bm_sad2_kernel", 1: "bm_sad_wta_kernel (generic)"}.get(matcher.last_kernel(), "?"),
            "bound": "int_alu", "achieved": ach, "peak": ip["iadd3_tiops"], "unit": "Tiop/s", "frac": ach / ip["iadd3_tiops"],
            "whole_step_frac": de * OPS_PER_DE["bm720"] / (step_ms * 1e-3) / 1e12 / ip["iadd3_tiops"],
            "traffic": (tr.get("dram_bytes_per_frame") or 0) * B or None,
            "traffic_note": tr.get("source"),
            "ops_per_de": OPS_PER_DE["bm720"], "kernel_ms_per_launch": k_ms, "frames_per_launch": B,
            "kernel_share_of_step": k_ms / step_ms,
            "peak_source": "rtdm_measure_int_peak on this GPU (dependent-free IADD3, lane-ops/s)", "int_peak_detail": ip,
            "hbm": {"achieved": hbm_ach, "peak": ctx.hbm_peak, "unit": "GB/s", "frac": hbm_ach / ctx.hbm_peak,
                    "peak_source": ctx.hbm_src, "algorithmic_bytes_per_frame": HBM_BYTES_PER_FRAME["bm720"]},
            "stage_ms_per_step": {k: v / stage_calls for k, v in stage_ms.items()},
        }
    return res


def bench_sgbm(ctx, K, frames, name):
    torch, rt = ctx.torch, ctx.rt
    from rtdm_b200 import sharding
    # 2 x 212 MB of cost volumes per frame; 60 frames = 4 rounds of the 15 thread-block clusters (one frame each, 9 CTAs)
    # a B200 keeps resident in the whole-height aggregation passes (rtdm_sgbm_batch_quantum)
    B = SGBM_BATCH
    Lh, Rh = frames[0][:B], frames[1][:B]
    dev, local = ctx.dev, ctx.local
    L, R = torch.from_numpy(Lh).to(dev), torch.from_numpy(Rh).to(dev)
    D = torch.empty((B, H, W), dtype=torch.int16, device=dev)
    matcher = rt.CUDASemiGlobalMatcher(SGBM_PARAMS["blockSize"], 0, ND, SGBM_PARAMS["uniquenessRatio"],
                                       SGBM_PARAMS["speckleWindowSize"], SGBM_PARAMS["speckleRange"],
                                       SGBM_PARAMS["disp12MaxDiff"], mode=SGBM_MODES[name],
                                       max_width=W, max_height=H, max_batch=B, device=local)
    st = torch.cuda.Stream(device=dev)

    def step():
        matcher.compute_device(B, L.data_ptr(), W, W * H, R.data_ptr(), W, W * H, W, H,
                               D.data_ptr(), W * 2, W * H * 2, st.cuda_stream)

    sampler = ClockSampler(local).start() if ctx.rank == 0 else None
    with torch.cuda.stream(st):
        for _ in range(max(2, min(ctx.args.warmup, 3))):
            step()
    launches_per_step = matcher.last_launches()
    Ks = max(2, min(K, 8))                                 # a 60-frame step lasts ~36 ms
    ctx.before_timed = lambda: matcher.set_profiling(True)
    ms, repeats = timed_region(ctx, step, st, Ks)
    nsteps = Ks * repeats
    clocks = sampler.stop() if sampler else None
    stage_ms, stage_calls = matcher.stage_times()
    matcher.set_profiling(False)
    fps, ms_max, _ = sharding.whole_job_throughput(B * nsteps, ms, 1.0, ctx.dist, dev)
    dev_out = D[[0, B - 1]].cpu().numpy()

    Lp, Rp = (torch.from_numpy(a).pin_memory().numpy() for a in (Lh, Rh))
    Dp = [torch.empty((B, H, W), dtype=torch.int16).pin_memory().numpy() for _ in range(2)]
    acc = [0]

    def e2e_run(n):
        for i in range(n):
            matcher.submit_batch(Lp, Rp, Dp[i & 1])
            if i > 0:
                matcher.wait_oldest()
                acc[0] += int(Dp[(i - 1) & 1][0, H // 2, W // 2])
        matcher.wait()
        acc[0] += int(Dp[(n - 1) & 1][0, H // 2, W // 2])

    e2e_run(2)
    e2e_steps = max(2, int(math.ceil(MIN_REGION_S * 1.05 / (ms_max / nsteps * 1e-3))))
    for d in Dp:
        d[:] = 0x5555
    ctx.barrier()
    t0 = time.perf_counter()
    e2e_run(e2e_steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_fps, e2e_ms_max, _ = sharding.whole_job_throughput(B * e2e_steps, e2e_s * 1e3, 1.0, ctx.dist, dev)
    res = {"workload": WORKLOAD_SGBM[name], "value": fps * mde_per_frame(), "unit": "Mde/s", "fps": fps,
           "ms_per_step": ms_max / nsteps, "steps": Ks, "repeats": repeats, "frames_per_step_per_gpu": B,
           "timed_region_s": ms_max * 1e-3, "clocks": clocks, "gpu_launches": launches_per_step * nsteps,
           "e2e": {"value": e2e_fps * mde_per_frame(), "unit": "Mde/s", "fps": e2e_fps, "h2d_bytes_per_step": B * 2 * W * H,
                   "d2h_bytes_per_step": B * W * H * 2, "steps": e2e_steps, "timed_region_s": e2e_ms_max * 1e-3,
                   "api": "rtdm_sgbm_submit_batch (2 batches in flight), results of batch i read during batch i+1 "
                          "(rtdm_sgbm_wait_oldest), pinned host buffers"},
           "_out": (Dp[(e2e_steps - 1) & 1], dev_out)}
    if ctx.rank == 0 and stage_calls:
        # the matching stage as a whole (planes, fused cost, horizontal paths, row sweeps, WTA, LR check): SURVEY.md 8(d)
        # rates it against the integer pipes (96 / 72 ops per de); the HBM figures sit beside it
        k_ms = stage_ms["matching"] / stage_calls
        ip = ctx.int_peak
        hbm_ach = B * HBM_BYTES_PER_FRAME[name] / (k_ms * 1e-3) / 1e9
        ach = B * W * H * ND * OPS_PER_DE[name] / (k_ms * 1e-3) / 1e12
        tr = ctx.traffic.get("sgbm_" + name, {})
        traffic = (tr.get("dram_bytes_per_frame") or 0) * B or None
        res["roofline"] = {
            "kernel": "sgbm matching stage (sgbm_cost_fused + sgbm_path4 first/last + sgbm_vpass cluster passes + sgbm_lr)",
            "bound": "int_alu", "achieved": ach, "peak": ip["iadd3_tiops"], "unit": "Tiop/s", "frac": ach / ip["iadd3_tiops"],
            "whole_step_frac": B * W * H * ND * OPS_PER_DE[name] / (ms / nsteps * 1e-3) / 1e12 / ip["iadd3_tiops"],
            "traffic": traffic, "traffic_note": tr.get("source"),
            "ops_per_de": OPS_PER_DE[name], "stage_ms_per_launch": k_ms, "frames_per_launch": B,
            "peak_source": "rtdm_measure_int_peak on this GPU (dependent-free IADD3, lane-ops/s)",
            "hbm": {"achieved": hbm_ach, "peak": ctx.hbm_peak, "unit": "GB/s", "frac": hbm_ach / ctx.hbm_peak, "peak_source": ctx.hbm_src,
                    "algorithmic_bytes_per_frame": HBM_BYTES_PER_FRAME[name],
                    "moved_gbs": traffic / (k_ms * 1e-3) / 1e9 if traffic else None,
                    "moved_frac_of_peak": traffic / (k_ms * 1e-3) / 1e9 / ctx.hbm_peak if traffic else None},
            "stage_ms_per_step": {k: v / stage_calls for k, v in stage_ms.items()},
        }
    del matcher
    return res


def bench_latency(ctx, iters=60):
    """Estimator::run's per-frame calls through the plugin API (estimator.cpp:45,54-56), one frame at a time, synchronous,
    host pointers; checked against cv2 on the fly."""
    torch, rt = ctx.torch, ctx.rt
    Ls, Rs, Ms = op_frames()
    rois = op_rois()
    Lp = [torch.from_numpy(a).pin_memory().numpy() for a in Ls]
    Rp = [torch.from_numpy(a).pin_memory().numpy() for a in Rs]
    out = torch.empty((OPH, OPW), dtype=torch.int16).pin_memory().numpy()
    bm = rt.CUDAMatcherKonolige(None, None, 31, 13, 0, 10, OPND, OPND, 10, 100, 32, 1, max_width=OPW, max_height=OPH, device=ctx.local)
    filt = rt.CUDAMorphologicalFilter(OPW, OPH, 8, device=ctx.local)
    tb = tf = 0.0
    checked = mism = 0
    for i in range(-5, iters):
        k = i % len(Ls)
        Lv, Rv = Lp[k][OPY:OPY + OPH, OPX:OPX + OPW], Rp[k][OPY:OPY + OPH, OPX:OPX + OPW]      # step = full image width
        filt.getVideoInBuffer()[:] = Ms[k]          # Estimator's inRange writes here (estimator.cpp:43); not part of the calls timed
        t0 = time.perf_counter()
        filt.run()                                  # morphFilter->run(filter_in, filter_out)
        t1 = time.perf_counter()
        bm.setROI1(rois[i % len(rois)])             # bm->setROI1(matching_roi)
        bm.compute(Lv, Rv, out)                     # bm->compute(left_rect, right_rect, left_disp)
        t2 = time.perf_counter()
        if i >= 0:
            tf += t1 - t0; tb += t2 - t1
        if i in (0, 1, 2, 3) and ctx.checker is not None:
            ref_d, ref_m = ctx.checker("op", (Ls[k][OPY:OPY + OPH, OPX:OPX + OPW], Rs[k][OPY:OPY + OPH, OPX:OPX + OPW], Ms[k], rois[i % len(rois)]))
            mism += int((ref_d != out).sum()) + int((ref_m != filt.getVideoOutBuffer()).sum())
            checked += 1
    # the reference's other matcher at the same operating point: SWSemiGlobalMatcher::compute (sgbm-sw.cpp:32-37; MODE_SGBM is
    # its literal default, P1 / P2 hard-coded), same strided views, one frame per call
    sg = rt.CUDASemiGlobalMatcher(SGBM_PARAMS["blockSize"], 0, OPND, SGBM_PARAMS["uniquenessRatio"], SGBM_PARAMS["speckleWindowSize"],
                                  SGBM_PARAMS["speckleRange"], SGBM_PARAMS["disp12MaxDiff"], mode=SGBM_MODES["mode_sgbm"],
                                  max_width=OPW, max_height=OPH, device=ctx.local)
    ts, sg_iters = 0.0, max(4, iters // 4)
    for i in range(-2, sg_iters):
        k = i % len(Ls)
        Lv, Rv = Lp[k][OPY:OPY + OPH, OPX:OPX + OPW], Rp[k][OPY:OPY + OPH, OPX:OPX + OPW]
        t0 = time.perf_counter()
        sg.compute(Lv, Rv, out)
        if i >= 0:
            ts += time.perf_counter() - t0
        if i == 0 and ctx.checker is not None:
            mism += int((ctx.checker("op_sgbm", (Ls[k][OPY:OPY + OPH, OPX:OPX + OPW], Rs[k][OPY:OPY + OPH, OPX:OPX + OPW])) != out).sum())
            checked += 1
    return {"ours": {"filter_ms": tf / iters * 1e3, "bm_ms": tb / iters * 1e3, "frame_ms": (tf + tb) / iters * 1e3, "frames": iters,
                     "api": "rtdm_morph_run + rtdm_bm_set_roi1 + rtdm_bm_compute, host pointers, synchronous",
                     "kernel": bm.last_kernel(), "sgbm_ms": ts / sg_iters * 1e3, "sgbm_frames": sg_iters,
                     "sgbm_api": f"rtdm_sgbm_compute, MODE_SGBM {OPW}x{OPH} nd={OPND} bs={SGBM_PARAMS['blockSize']}, host pointers, synchronous"},
            "_parity": (checked, mism)}


def make_checker():
    """-> fn(kind, payload) computing the reference result with cv2 (or the C oracle when cv2 is missing)."""
    from oracle import cv2_ref
    if cv2_ref.have_cv2():
        import cv2
        cv2.setNumThreads(os.cpu_count() or 1)
        bm = cv2_ref.make_bm(**BM_PARAMS)
        bm_op = cv2_ref.make_bm(**dict(BM_PARAMS, numDisparities=OPND))
        sg = {n: cv2_ref.make_sgbm(mode=m, **SGBM_PARAMS) for n, m in SGBM_MODES.items()}
        sg_op = cv2_ref.make_sgbm(mode=SGBM_MODES["mode_sgbm"], **dict(SGBM_PARAMS, numDisparities=OPND))

        def check(kind, x):
            if kind == "bm":
                return bm.compute(x[0], x[1])
            if kind == "mask":
                return cv2_ref.morph_open_close(x)
            if kind == "op":
                # cv2 never writes the rows of stripes that lie wholly outside the valid rectangle (a fresh output Mat
                # holds heap garbage there, run to run different); the reference reuses one left_disp Mat, so hand cv2
                # an output that already holds FILTERED (-16), which is what this library writes there
                bm_op.setROI1(x[3])
                out = np.full(x[0].shape, -16, np.int16)
                return bm_op.compute(x[0], x[1], out), cv2_ref.morph_open_close(x[2])
            if kind == "op_sgbm":
                return sg_op.compute(x[0], x[1])
            return sg[kind].compute(x[0], x[1])
        return check, f"cv2 {cv2.__version__}"
    from oracle import oracle

    def check(kind, x):
        if kind == "bm":
            return oracle.bm_compute(x[0], x[1], oracle.make_params(**BM_PARAMS))
        if kind == "mask":
            return oracle.morph_open_close(x)
        if kind == "op":
            return (oracle.bm_compute(np.ascontiguousarray(x[0]), np.ascontiguousarray(x[1]),
                                      oracle.make_params(**dict(BM_PARAMS, numDisparities=OPND, roi1=x[3]))), oracle.morph_open_close(x[2]))
        if kind == "op_sgbm":
            return oracle.sgbm_compute(np.ascontiguousarray(x[0]), np.ascontiguousarray(x[1]),
                                       oracle.make_params(P1=600, P2=2400, preFilterCap=0, mode=SGBM_MODES["mode_sgbm"], **dict(SGBM_PARAMS, numDisparities=OPND)))
        return oracle.sgbm_compute(x[0], x[1], oracle.make_params(P1=600, P2=2400, preFilterCap=0, mode=SGBM_MODES[kind], **SGBM_PARAMS))
    return check, "oracle/stereo_oracle.c"


def run_ours(args):
    import torch
    import rtdm_b200 as rt

    ctx = Ctx()
    ctx.torch, ctx.rt, ctx.args = torch, rt, args
    ctx.rank = rank = int(os.environ.get("RANK", "0"))
    ctx.world = world = int(os.environ.get("WORLD_SIZE", "1"))
    ctx.local = local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or rt.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    ctx.dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        ctx.dist = dist
    ctx.dev = torch.device("cuda", local)

    def barrier():
        if ctx.dist is not None:
            ctx.dist.barrier()
        torch.cuda.synchronize()
    ctx.barrier = barrier
    ctx.before_timed = lambda: None
    ctx.traffic = load_traffic()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    ctx.hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    ctx.hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    ctx.int_peak = rt.measure_int_peak(local) if rank == 0 else None
    ctx.checker, checker_name = (make_checker() if rank == 0 and not args.no_check else (None, None))
    ctx.link_gbs = link_ceiling(ctx)

    K = max(1, args.steps)
    wl = args.workload
    # every rank works on its own distinct frames (seeds 1000 + rank * batch + i)
    frames = make_frames(BM_BATCH, 1000 + rank * BM_BATCH)
    bm = bench_bm(ctx, K, frames) if wl in ("all", "bm720") else None
    sg = {}
    if wl in ("all", "sgbm720"):
        for name in SGBM_MODES:
            sg[name] = bench_sgbm(ctx, K, frames, name)
    lat = bench_latency(ctx) if (wl == "all" and rank == 0 and not args.no_latency) else None
    barrier()

    failure = None
    if rank == 0:
        # ---- parity of what the timed regions produced (e2e output buffers + device-side outputs) ---------
        parity = None
        if ctx.checker is not None:
            parity = {"checker": checker_name, "mismatching_pixels": 0, "by_leg": {}}
            Lh, Rh, Mh = frames
            if bm is not None:
                Dp, MOp, (dsel, msel) = bm["_out"]
                idx = [0, BM_BATCH // 3, BM_BATCH // 2, (2 * BM_BATCH) // 3, BM_BATCH - 1]
                bad = 0
                for i in idx:
                    ref = ctx.checker("bm", (Lh[i], Rh[i]))
                    bad += int((ref != Dp[i]).sum())
                    if i in (0, BM_BATCH // 2, BM_BATCH - 1):
                        bad += int((ref != dsel[[0, BM_BATCH // 2, BM_BATCH - 1].index(i)]).sum())
                refm = ctx.checker("mask", Mh[0])
                bad += int((refm != MOp[0]).sum()) + int((refm != msel[0]).sum())
                refm = ctx.checker("mask", Mh[BM_BATCH - 1])
                bad += int((refm != MOp[BM_BATCH - 1]).sum())
                parity["bm"] = len(idx); parity["masks"] = 2
                parity["mismatching_pixels"] += bad
                parity["by_leg"]["bm"] = bad
            for name, r in sg.items():
                Dp, dsel = r["_out"]
                idx = [0, SGBM_BATCH // 3, (2 * SGBM_BATCH) // 3, SGBM_BATCH - 1]
                bad = 0
                for i in idx:
                    ref = ctx.checker(name, (Lh[i], Rh[i]))
                    bad += int((ref != Dp[i]).sum())
                    if i in (0, SGBM_BATCH - 1):
                        bad += int((ref != dsel[0 if i == 0 else 1]).sum())
                parity["sgbm_" + name] = len(idx)
                parity["mismatching_pixels"] += bad
                parity["by_leg"]["sgbm_" + name] = bad
            if sg:
                parity["sgbm"] = sum(parity["sgbm_" + n] for n in sg)
            if lat is not None:
                parity["latency_frames"] = lat["_parity"][0]
                parity["mismatching_pixels"] += lat["_parity"][1]
                parity["by_leg"]["latency"] = lat["_parity"][1]
        # ---- CPU baselines on bounded samples (rank 0, N=1 only) ------------------------------------------
        if world == 1 and not args.no_cpu:
            sub = (frames[0][:16], frames[1][:16], frames[2][:16])
            if bm is not None:
                bm["cpu_baseline"] = cpu_baseline_obj("bm720", sub, 240, 16)
            for name, r in sg.items():
                r["cpu_baseline"] = cpu_baseline_obj(name, (sub[0][:4], sub[1][:4], None), 8, 0)
                r["cpu_baseline"]["note"] = "MODE_SGBM / MODE_HH are single-threaded inside OpenCV (SURVEY.md section 6)"
            if lat is not None:
                lat["reference"] = cpu_latency()
        for r in [bm] + list(sg.values()) + [lat]:
            if r is not None:
                for k in [k for k in r if k.startswith("_")]:
                    del r[k]
        top = bm if bm is not None else sg["mode_hh"]
        line = {
            "metric": "Mde/s", "value": top["value"], "unit": "Mde/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": top["ms_per_step"], "repeats": top["repeats"], "timed_region_s": top["timed_region_s"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/u16 integer", "data": "synthetic", "fps": top["fps"],
            "config": config_for(world),
            "e2e": top["e2e"], "gpu_launches": top["gpu_launches"] + (sum(r["gpu_launches"] for r in sg.values()) if bm is not None else 0),
            "clocks": top["clocks"], "roofline": top.get("roofline"), "cpu_baseline": top.get("cpu_baseline"),
            "parity_checked": parity,
        }
        if sg:
            line["sgbm"] = sg
        if lat is not None:
            line["latency"] = lat
        print(json.dumps(line), flush=True)
        if parity is not None and parity["mismatching_pixels"] != 0:
            failure = f"bench.py: PARITY FAILURE -- {parity['mismatching_pixels']} pixels of the timed outputs differ from {checker_name}"
    if ctx.dist is not None:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()
    if failure:
        raise SystemExit(failure)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="all", choices=["all", "bm720", "sgbm720"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-check", action="store_true", help="skip the parity check of the timed outputs")
    ap.add_argument("--no-latency", action="store_true", help="skip the per-frame latency leg")
    ap.add_argument("--min-region-s", type=float, default=MIN_REGION_S,
                    help="minimum length of every timed region (profiling runs under ncu pass 0: K steps exactly)")
    args = ap.parse_args()
    globals()["MIN_REGION_S"] = max(0.0, args.min_region_s)
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
