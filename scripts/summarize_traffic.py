"""Per-kernel totals of an ncu launch list that carries gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum."""
import collections, csv, re, sys

SCALE_T = {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0}
SCALE_B = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
with open(sys.argv[1]) as fh:
    lines = [l for l in fh if l.startswith('"')]
per = collections.OrderedDict()
for r in csv.DictReader(lines):
    name = re.sub(r"^.*::", "", re.sub(r"\(.*", "", r["Kernel Name"]))
    name = re.sub(r"<.*", "", name)
    k = per.setdefault((r["ID"], name), {})
    v, u = float(r["Metric Value"].replace(",", "")), r["Metric Unit"]
    m = r["Metric Name"]
    k[m] = v * (SCALE_T.get(u, 1e-6) if "time" in m else SCALE_B.get(u, 1.0))
tot = collections.OrderedDict()
for (_, name), k in per.items():
    t = tot.setdefault(name, [0, 0.0, 0.0, 0.0])
    t[0] += 1; t[1] += k.get("gpu__time_duration.sum", 0); t[2] += k.get("dram__bytes_read.sum", 0); t[3] += k.get("dram__bytes_write.sum", 0)
T = sum(t[1] for t in tot.values()); R = sum(t[2] for t in tot.values()); W = sum(t[3] for t in tot.values())
print(f"launches={len(per)} time={T:.3f} ms dram_read={R/1e6:.1f} MB dram_write={W/1e6:.1f} MB traffic={(R+W)/1e6:.1f} MB")
for n, (c, t, r, w) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"  {n[:40]:40s} n={c:4d} time={t:8.3f} ms share={t/T:6.1%} read={r/1e6:9.1f} MB write={w/1e6:9.1f} MB  {((r+w)/1e9)/(t/1e3) if t else 0:7.1f} GB/s")
