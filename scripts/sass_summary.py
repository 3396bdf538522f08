"""Instruction-mix summary of the shipped cubins (cuobjdump -sass): per kernel, the SASS mnemonics that show how memory and
synchronisation are done (bulk / asynchronous copies, vector accesses, atomics, sleeps) next to the arithmetic it runs.
    python scripts/sass_summary.py tfmpc_b200/lib/libtfmpc_b200.so > profiles/r02_sass_summary.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1]
res = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
usage = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
WATCH = ["UBLKCP", "UTMALDG", "UTCMMA", "LDGSTS", "SYNCS", "LDG.E.128", "STG.E.128", "LDS.128", "STS.128", "ATOMG", "ATOM", "RED", "NANOSLEEP",
         "SHFL", "VOTE", "MUFU", "FFMA", "DFMA", "LDL", "STL", "WARPSYNC", "CCTL", "MEMBAR", "FENCE"]
print(f"# {lib}: cuobjdump -sass / -res-usage, static counts of selected mnemonics per kernel")
archs = sorted(set(re.findall(r"arch = (sm_\w+)", res)))
print("# architectures in the fat binary:", ", ".join(archs))
regs = {}
fn = None
for line in usage.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        fn = m.group(1)
        continue
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
    if m and fn:
        regs[fn] = tuple(int(v) for v in m.groups())
cur, counts, total = None, None, 0
rows = []
for line in res.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        if cur:
            rows.append((cur, total, counts))
        cur, counts, total = m.group(1), collections.Counter(), 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        total += 1
        for w in WATCH:
            base = w.split(".")[0]
            if "128" in w:
                hit = op.startswith(base + ".") and ".128" in op
            else:
                hit = op == w or op.startswith(w + ".")
            if hit:
                counts[w] += 1
if cur:
    rows.append((cur, total, counts))
dem = subprocess.run(["cu++filt"] + [r[0] for r in rows], capture_output=True, text=True).stdout.splitlines()
for (name, total, counts), d in zip(rows, dem):
    short = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", d)
    short = re.sub(r"^void ", "", short)
    short = re.sub(r">\(.*$", ">", short)[:64]
    r = regs.get(name)
    ru = f"regs={r[0]:3d} stack={r[1]:4d} smem={r[2]:6d}" if r else ""
    keep = {k: v for k, v in counts.items() if v}
    print(f"{short:66s} instr={total:6d} {ru}  " + " ".join(f"{k}={v}" for k, v in sorted(keep.items())))
