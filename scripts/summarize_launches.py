"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and the per-launch series."""
import csv, sys, re, collections
rows = []
with open(sys.argv[1]) as fh:
    lines = [l for l in fh if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit, 1e-3)
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"^.*::", "", name)
    rows.append((name, v * scale))
tot = collections.OrderedDict()
for n, v in rows:
    t = tot.setdefault(n, [0, 0.0]); t[0] += 1; t[1] += v
total = sum(v for _, v in rows)
print(f"launches={len(rows)} total={total/1e3:.3f} ms")
for n, (c, v) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"  {n[:60]:60s} n={c:4d} total={v/1e3:9.3f} ms  share={v/total:6.1%}  mean={v/c:9.1f} us")
if len(sys.argv) > 2:
    pat = sys.argv[2]
    series = [v for n, v in rows if pat in n]
    print(pat, "series (us):", " ".join(f"{v:.0f}" for v in series))
