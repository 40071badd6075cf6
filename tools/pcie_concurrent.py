"""Host-link ceiling with 1 .. N GPUs busy at once: every process (one per GPU) copies pinned 64 MB blocks host->device and
device->host SIMULTANEOUSLY (two streams), all processes start together.  Prints one CSV row per N:
n_gpus, per-GPU H2D GB/s (mean / min), per-GPU D2H GB/s (mean / min), aggregate GB/s per direction.
  python tools/pcie_concurrent.py [max_gpus] [--affinity]   (profiles/r02_pcie_concurrent.csv is its output on the 8-GPU box)"""
import os, sys, time
import multiprocessing as mp


def worker(rank, n, barrier, q, affinity):
    import torch
    if affinity:
        ncpu = os.cpu_count() or 1
        per = max(1, ncpu // n)
        os.sched_setaffinity(0, set(range(rank * per, min(ncpu, (rank + 1) * per))))
    torch.cuda.set_device(rank)
    nb = 64 << 20
    h1 = torch.empty(nb, dtype=torch.uint8).pin_memory(); d1 = torch.empty(nb, dtype=torch.uint8, device="cuda")
    h2 = torch.empty(nb, dtype=torch.uint8).pin_memory(); d2 = torch.zeros(nb, dtype=torch.uint8, device="cuda")
    h1.fill_(1)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

    def go(k, timed):
        if timed:
            e[0].record(s1); e[2].record(s2)
        for _ in range(k):
            with torch.cuda.stream(s1):
                d1.copy_(h1, non_blocking=True)
            with torch.cuda.stream(s2):
                h2.copy_(d2, non_blocking=True)
        if timed:
            e[1].record(s1); e[3].record(s2)
    go(3, False)
    torch.cuda.synchronize()
    barrier.wait()
    K = 24
    t0 = time.perf_counter()
    go(K, True)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    q.put((rank, K * nb / (e[0].elapsed_time(e[1]) * 1e-3) / 1e9, K * nb / (e[2].elapsed_time(e[3]) * 1e-3) / 1e9, K * nb / wall / 1e9))


def main():
    import torch
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    affinity = "--affinity" in sys.argv
    nmax = int(args[0]) if args else torch.cuda.device_count()
    ctx = mp.get_context("spawn")
    print("n_gpus,affinity,h2d_gbs_mean,h2d_gbs_min,d2h_gbs_mean,d2h_gbs_min,aggregate_gbs_per_direction_wall", flush=True)
    n = 1
    while n <= nmax:
        barrier, q = ctx.Barrier(n), ctx.Queue()
        ps = [ctx.Process(target=worker, args=(r, n, barrier, q, affinity)) for r in range(n)]
        for p in ps: p.start()
        res = [q.get(timeout=300) for _ in range(n)]
        for p in ps: p.join()
        h2d = [r[1] for r in res]; d2h = [r[2] for r in res]
        print(f"{n},{int(affinity)},{sum(h2d)/n:.1f},{min(h2d):.1f},{sum(d2h)/n:.1f},{min(d2h):.1f},{sum(r[3] for r in res):.1f}", flush=True)
        n *= 2


if __name__ == "__main__":
    main()
