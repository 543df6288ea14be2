"""Print the SASS of one kernel from an ncu report with per-instruction execution counts.
usage: ncu_sass.py report.ncu-rep kernel-regex [min_exec]"""
import csv, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
mn = int(sys.argv[3]) if len(sys.argv) > 3 else 0
out = subprocess.run(["ncu", "-i", rep, "-k", "regex:" + kre, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hd = None
n = 0
tot = 0
seen = 0
for r in rows:
    if r and r[0] == "Kernel Name":
        seen += 1
        if seen > 1:
            break  # the same kernel captured more than once: first capture only
    if len(r) > 5 and r[0] == "Address":
        hd = r
        ie, it, sm = hd.index("Instructions Executed"), hd.index("Avg. Threads Executed"), hd.index("# Samples")
        continue
    if hd and len(r) == len(hd):
        try:
            e = int(r[ie])
        except ValueError:
            continue
        n += 1
        tot += e
        if e >= mn:
            print("%5d %9d %5s %5s  %s" % (n, e, r[it], r[sm], r[1].strip()))
print("instructions", n, "executed", tot)
