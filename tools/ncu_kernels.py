"""One summary CSV per kernel from a multi-kernel `ncu --set full` report:
   python tools/ncu_kernels.py gpurun_out/<rep>.ncu-rep profiles/r02_full_<tag> [frames_per_launch]
Each file: duration, DRAM bytes, launch shape, issue-slot / pipe utilisation, shared-memory conflicts, the stall reasons
ranked (warps stalled per issue-active cycle).  Only the first launch of every distinct kernel is kept."""
import csv, os, re, subprocess, sys

KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__grid_size', 'launch__block_size',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static', 'launch__waves_per_multiprocessor',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__warps_eligible.avg.per_cycle_active']


def main(rep, prefix, frames=None):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h, units = rows[0], rows[1]
    ki = h.index('Kernel Name')
    stalls = [i for i, n in enumerate(h) if re.match(r'smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio$', n) and 'not_issued' not in n]
    seen = set()
    for r in rows[2:]:
        name = r[ki]
        short = re.sub(r'[^A-Za-z0-9]+', '_', re.sub(r'\(.*', '', name.replace('void ', '').replace('unnamed>::', ''))).strip('_')
        if short in seen:
            continue
        seen.add(short)
        out = [f"# ncu --set full --clock-control none, first captured launch of this kernel; report {os.path.basename(rep)} (scratch, not committed)"]
        if frames:
            out.append(f"# one launch = {frames} frames")
        out.append(f"Kernel Name,\"{name}\",")
        for n in KEEP:
            if n in h:
                out.append(f"{n},{r[h.index(n)]},{units[h.index(n)]}")
        ranked = sorted(((float(r[i].replace(',', '') or 0), h[i]) for i in stalls), reverse=True)
        out.append("# stall reasons, warps per issue-active cycle, ranked")
        for v, n in ranked[:6]:
            out.append(f"{n.replace('smsp__average_warps_issue_stalled_', 'stall_').replace('_per_issue_active.ratio', '')},{v:.3f},warps/issue")
        path = f"{prefix}_{short}.csv"
        open(path, "w").write("\n".join(out) + "\n")
        d = {n: r[h.index(n)] for n in KEEP if n in h}
        print(f"{short}: {d.get('gpu__time_duration.sum')} {units[h.index('gpu__time_duration.sum')]}, dram {d.get('dram__bytes_read.sum')}+{d.get('dram__bytes_write.sum')} {units[h.index('dram__bytes_read.sum')]}, "
              f"issue {d.get('smsp__issue_active.avg.pct_of_peak_sustained_active')}%, dram% {d.get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')}, regs {d.get('launch__registers_per_thread')}, "
              f"top stalls: " + ", ".join(f"{n.split('stalled_')[1].split('_per_')[0]} {v:.2f}" for v, n in ranked[:3]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
