"""Summarise an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file X.csv <command>):
per kernel count, total and average duration, share of the listed time.
usage: python scripts/launch_summary.py X.csv "header comment" > profiles/X_summary.txt"""
import collections, csv, re, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hd = rows[0]
ki, vi, ui = hd.index("Kernel Name"), hd.index("Metric Value"), hd.index("Metric Unit")
acc = collections.OrderedDict()
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ki])
    v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
    c = acc.setdefault(name, [0, 0.0])
    c[0] += 1
    c[1] += v
tot = sum(c[1] for c in acc.values())
if len(sys.argv) > 2:
    print("# " + sys.argv[2])
print("# (cold-cache, serialised launches: shares matter, absolutes do not)")
print("%-62s %6s %12s %8s %7s" % ("kernel", "count", "total_us", "avg_us", "share"))
for name, (n, t) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print("%-62s %6d %12.1f %8.1f %6.1f%%" % (name[:62], n, t, t / n, 100 * t / tot))
print("%-62s %6d %12.1f" % ("total", sum(c[0] for c in acc.values()), tot))
