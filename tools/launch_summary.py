"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: count, total, share and mean per kernel.
Usage: launch_summary.py launches.csv ["command line that was profiled"]"""
import collections, csv, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')) if len(r) > 14]
h = rows[0]
kn, mv = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[kn].split("(")[0][:70]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[mv].replace(",", "")) / 1e6
tot = sum(a[1] for a in agg.values())
if len(sys.argv) > 2:
    print(sys.argv[2])
print("(per-launch times are cold-cache and serialised under ncu: compare SHARES, not absolute times)\n")
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:72s} n={n:4d} total={t:9.3f} ms share={100 * t / tot:5.1f}%  mean={1e3 * t / n:9.1f} us")
