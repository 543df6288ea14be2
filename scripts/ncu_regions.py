"""Aggregate ncu per-line instruction counts into user-given line regions: file:lo-hi:name ..."""
import csv, os, shlex, subprocess, sys
EXTRA = shlex.split(os.environ.get("NCU_EXTRA", ""))  # e.g. NCU_EXTRA="-k regex:parse_kernel_v4"
rep = sys.argv[1]
regions = []
for a in sys.argv[2:]:
    f, rng, name = a.split(":")
    lo, hi = rng.split("-")
    regions.append((f, int(lo), int(hi), name))
src = subprocess.run(["ncu", "-i", rep, *EXTRA, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
cur, hd, items = None, None, []
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]; continue
    if len(r) > 5 and r[0] == 'Line No':
        hd = r; continue
    if hd and len(r) == len(hd) and r[2] == '-':
        try:
            n = int(r[hd.index('Instructions Executed')]); s = int(r[hd.index('# Samples')]); t = int(r[hd.index('Thread Instructions Executed')])
        except ValueError:
            continue
        items.append((n, s, cur, int(r[0]), t))
tot = sum(i[0] for i in items); tots = sum(i[1] for i in items)
print("total warp-inst %d samples %d" % (tot, tots))
used = set()
for f, lo, hi, name in regions:
    sel = [i for i in items if i[2] == f and lo <= i[3] <= hi]
    for i in sel: used.add((i[2], i[3]))
    n = sum(i[0] for i in sel); s = sum(i[1] for i in sel); t = sum(i[4] for i in sel)
    print("%-34s inst %5.1f%% (%7.1fM) samp %5.1f%% thr %4.1f" % (name, 100 * n / tot, n / 1e6, 100 * s / tots, t / max(n, 1)))
rest = [i for i in items if (i[2], i[3]) not in used]
byfile = {}
for i in rest:
    byfile.setdefault(i[2], [0, 0])
    byfile[i[2]][0] += i[0]; byfile[i[2]][1] += i[1]
for f, (n, s) in byfile.items():
    print("rest %-29s inst %5.1f%% (%7.1fM) samp %5.1f%%" % (f, 100 * n / tot, n / 1e6, 100 * s / tots))
