"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
Usage: python tools/ncu_launches.py gpurun_out/launches.csv > profiles/rNN_launches_summary.txt"""
import csv, re, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 14 and r[0].isdigit()]
tot = collections.OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("<unnamed>::", "").replace("void ", "").strip()
    t = tot.setdefault(name, [0, 0.0])
    t[0] += 1; t[1] += float(r[14]) / 1000.0
allus = sum(v[1] for v in tot.values())
print("%-28s %8s %12s %10s %7s" % ("kernel", "launches", "total_us", "avg_us", "share"))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-28s %8d %12.1f %10.2f %6.1f%%" % (k, v[0], v[1], v[1] / v[0], 100 * v[1] / allus))
