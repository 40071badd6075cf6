"""Development aid: per-instruction stall samples of one kernel from an ncu report's source page.
   python tools/ncu_hot.py <rep> [top]   -> BAR / SYNCS instructions with their stall samples, the top-N sampled instructions,
   and totals per stall reason."""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
h = rows[1]; data = rows[2:]
iS, iN, iE = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
st = [(i, n) for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
tot = collections.Counter(); total = 0
for r in data:
    total += int(r[iN] or 0)
    for i, n in st: tot[n] += int(r[i] or 0)
print("total samples", total, {n: v for n, v in tot.most_common(8)})
execd = sum(int(r[iE] or 0) for r in data); print("warp instructions executed", execd)
print("--- barrier / mbarrier instructions")
for k, r in enumerate(data):
    s = r[iS]
    if "BAR." in s or "SYNCS" in s or "UBLKCP" in s or "UTMALDG" in s:
        d = {n: int(r[i] or 0) for i, n in st if int(r[i] or 0)}
        print(k, s.strip()[:60], "samples", r[iN], "exec", r[iE], d)
print("--- top sampled instructions")
for k, r in sorted(enumerate(data), key=lambda t: -int(t[1][iN] or 0))[:top]:
    d = {n: int(r[i] or 0) for i, n in st if int(r[i] or 0) > 0.1 * int(r[iN] or 1)}
    print(k, r[iS].strip()[:70], "samples", r[iN], "exec", r[iE], d)
