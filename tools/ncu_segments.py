"""Development aid: splits the SASS page of an ncu report (ncu -i X.ncu-rep --page source --csv --print-source sass)
at barrier instructions and prints instructions executed / stall samples per segment."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[1], rows[2:]
iS, iI, iW = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
cols = ['stall_barrier', 'stall_short_sb', 'stall_long_sb', 'stall_wait', 'stall_math', 'stall_mio', 'stall_not_selected', 'stall_selected', 'stall_branch_resolving', 'stall_no_inst', 'stall_dispatch', 'stall_lg']
idx = [hdr.index(c) for c in cols]
cuts = [0] + [n for n, r in enumerate(data) if 'BAR.' in r[iS] or 'EXIT' in r[iS]] + [len(data)]
tot = sum(int(r[iI] or 0) for r in data); tots = sum(int(r[iW] or 0) for r in data)
print("total inst", tot, "samples", tots)
for a, b in zip(cuts[:-1], cuts[1:]):
    seg = data[a:b + 1] if b < len(data) else data[a:b]
    n = sum(int(r[iI] or 0) for r in seg); s = sum(int(r[iW] or 0) for r in seg)
    if n == 0: continue
    ex = int(data[b][iI] or 0) if b < len(data) else 0
    d = {c: sum(int(r[i] or 0) for r in seg) for c, i in zip(cols, idx)}
    print(f"[{a:4d},{b:4d}] {data[b][iS].strip()[:38] if b < len(data) else 'end':38s} inst {n:10d} per-exec {n / max(ex, 1):7.1f} samples {s:6d}", {k[6:]: v for k, v in d.items() if v > s * 0.06})
