"""Turns ncu output under gpurun_out/ into the small CSV summaries committed under profiles/."""
import csv, collections, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

def launches(src, dst, title, per_launch_note):
    rows = list(csv.reader(l for l in open(os.path.join(G, src)) if l.startswith('"')))
    h = rows[0]; ki, vi, mi = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Name')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        agg.setdefault(r[ki], collections.OrderedDict()).setdefault(r[mi], []).append(float(r[vi].replace(',', '')))
    out = [f"# {title}", f"# {per_launch_note}",
           "# per-launch times are cold-cache and serialised under ncu: compare SHARES, not absolutes",
           "kernel,launches,avg_us,dram_read_MB,dram_write_MB,share_of_listed_time_pct"]
    tot = sum(sum(m.get('gpu__time_duration.sum', [0])) for k, m in agg.items() if 'intpeak' not in k)
    for k, m in agg.items():
        if 'intpeak' in k: continue
        t = m.get('gpu__time_duration.sum', [0]); rd = m.get('dram__bytes_read.sum', [0]); wr = m.get('dram__bytes_write.sum', [0])
        def mb(v, unit_guess):  # ncu prints bytes in varying units in csv; values here are already numeric in the unit column's scale
            return v
        out.append(f"\"{k}\",{len(t)},{sum(t)/len(t)/1e3:.1f},{sum(rd)/len(rd):.1f},{sum(wr)/len(wr):.1f},{sum(t)/tot*100:.1f}")
    open(os.path.join(P, dst), "w").write("\n".join(out) + "\n")
    print("\n".join(out))

def raw_summary(rep, dst):
    keep = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','launch__grid_size','launch__block_size','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__waves_per_multiprocessor','smsp__inst_executed.sum','smsp__inst_executed.max','smsp__inst_executed.min','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__issue_active.max.pct_of_peak_sustained_active','smsp__issue_active.min.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','dram__cycles_active.avg.pct_of_peak_sustained_elapsed','dram__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__warps_eligible.avg.per_cycle_active']
    txt = subprocess.run(["ncu", "-i", os.path.join(G, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines())); h = rows[0]
    out = [f"# ncu --set full --clock-control none, one launch; source: gpurun_out/{rep} (not committed, {os.path.getsize(os.path.join(G, rep))>>20} MiB)"]
    out.append(f"Kernel Name,{rows[2][h.index('Kernel Name')]},")
    for i, n in enumerate(h):
        if n in keep: out.append(f"{n},{rows[2][i]},{rows[1][i]}")
    open(os.path.join(P, dst), "w").write("\n".join(out) + "\n")
    print("\n".join(out[:40]))

if __name__ == "__main__":
    launches("launches_r1d.csv", "r01_launches_D_bm_fast_kernel.csv",
             "ncu launch list, round 1 capture D (fast BM kernel + binary filter fast path): python bench.py --steps 2 --warmup 3 --batch 16 --no-cpu",
             "one launch = 16 frames 1280x720 nd=128; dram columns are in the unit ncu printed (MB for the large kernels, KB/B for tiny ones)")
    launches("launches_sgbm_r1d.csv", "r01_launches_D_sgbm.csv",
             "ncu launch list, round 1 capture D (SGBM MODE_HH): python bench.py --workload sgbm720 --steps 1 --warmup 3 --batch 4 --no-cpu",
             "one launch = 4 frames 1280x720 nd=128; 8 path launches per step")
    raw_summary("prof_bm_r1d.ncu-rep", "r01_prof_bm_r1d_summary.csv")
    raw_summary("prof_sgbm_path_r1d.ncu-rep", "r01_prof_sgbm_path_r1d_summary.csv")
