"""Turns ncu output under gpurun_out/ into the small CSV summaries committed under profiles/."""
import csv, collections, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

def launches(src, dst, title, per_launch_note):
    rows = list(csv.reader(l for l in open(os.path.join(G, src)) if l.startswith('"')))
    h = rows[0]; ki, vi, mi, ui, ii = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Name'), h.index('Metric Unit'), h.index('ID')
    scale = {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3, 'ns': 1e-3, 'nsecond': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3, 'msecond': 1e3}
    agg = collections.OrderedDict()
    for r in rows[1:]:
        agg.setdefault(r[ki], collections.OrderedDict()).setdefault(r[mi], []).append(float(r[vi].replace(',', '')) * scale.get(r[ui], 1.0))
    out = [f"# {title}", f"# {per_launch_note}",
           "# per-launch times are cold-cache and serialised under ncu: compare SHARES, not absolutes",
           "kernel,launches,avg_us,dram_read_MB_per_launch,dram_write_MB_per_launch,share_of_listed_time_pct"]
    tot = sum(sum(m.get('gpu__time_duration.sum', [0])) for k, m in agg.items() if 'intpeak' not in k)
    for k, m in agg.items():
        if 'intpeak' in k: continue
        t = m.get('gpu__time_duration.sum', [0]); rd = m.get('dram__bytes_read.sum', [0]); wr = m.get('dram__bytes_write.sum', [0])
        out.append(f"\"{k}\",{len(t)},{sum(t)/len(t):.1f},{sum(rd)/len(rd):.1f},{sum(wr)/len(wr):.1f},{sum(t)/tot*100:.1f}")
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
    # usage: summarize_profiles.py launches <src.csv> <dst.csv> <title> <note>  |  raw <rep> <dst.csv>
    if sys.argv[1] == "launches":
        launches(*sys.argv[2:6])
    else:
        raw_summary(*sys.argv[2:4])
