"""Writes the SASS of the dominant kernels (from the in-tree librtdm_b200.so) and an instruction-mix table
under profiles/ (cuobjdump -sass; sm_100a)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "rt-depth-map_b200", "librtdm_b200.so")
OUT = os.path.join(ROOT, "profiles")
ROUND = "r02"
KERNELS = {                         # file tag -> substring of the mangled name
    "bm_sad4_h6_d128": "bm_sad4_kernelILi6ELi16E",
    "bm_sad3_h6_d128": "bm_sad3_kernelILi6ELi16ENS0_9ShapeWide",
    "bm_sad2_h6": "bm_sad2_kernelILi6ELi1ELi192ELb0ELi2E",
    "sgbm_vpass_d128": "sgbm_vpass_kernelILi16ELi4ELi1024ELi2ELb1E",
    "sgbm_vpass_d192": "sgbm_vpass_kernelILi32ELi3ELi1024ELi4ELb1E",
    "sgbm_sweep_d128": "sgbm_sweep_kernelILi16ELb1E",
    "sgbm_cost_fused_bs5": "sgbm_cost_fused_kernelILi5E",
    "sgbm_path4_d128_last": "sgbm_path4_kernelILi16ELi4ELi2E",
    "post_row8_validate": "post_row8_kernelILb1E",
}
txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
chunks = re.split(r"(?=\t\tFunction : )", txt)
rows = []
for tag, key in KERNELS.items():
    body = next(c for c in chunks if c.startswith("\t\tFunction : ") and key in c.split("\n", 1)[0])
    lines = [re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l) for l in body.splitlines() if not re.match(r"^\s*/\* 0x[0-9a-f]+ \*/\s*$", l)]
    with open(os.path.join(OUT, f"{ROUND}_sass_{tag}.txt"), "w") as f:
        f.write(f"# cuobjdump -sass librtdm_b200.so, function containing '{key}' (sm_100a, encodings stripped)\n")
        f.write("\n".join(lines) + "\n")
    ops = collections.Counter()
    for l in lines:
        m = re.match(r"^\s*/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", l)
        if m:
            ops[m.group(1)] += 1
    n = sum(ops.values())
    rows.append((tag, n, ops))
    print(tag, n, ops.most_common(8))
with open(os.path.join(OUT, f"{ROUND}_sass_instruction_mix.csv"), "w") as f:
    f.write("# static SASS instruction mix of the dominant kernels (counts over the whole function, all unrolled variants)\n")
    f.write("kernel,total,top opcodes (count)\n")
    for tag, n, ops in rows:
        f.write(f"{tag},{n},\"" + ", ".join(f"{k} {v}" for k, v in ops.most_common(14)) + "\"\n")
