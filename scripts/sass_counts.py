"""Per-kernel SASS mnemonic counts of the shipped library (static, from cuobjdump): which kernels use
the 1-D TMA bulk copies (UBLKCP), mbarriers (SYNCS), IDP.4A, global atomics / reductions ...
usage: python scripts/sass_counts.py [lib.so] > profiles/rNN_sass_counts.txt"""
import collections, os, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         "kata-xpu-device-plugin_b200", "lib", "libkxpu.so")
PREFIXES = ["UBLKCP", "SYNCS", "ATOMG", "ATOMS", "RED", "IDP.4A", "LDS.128", "LDG.E.128", "STG.E.128", "MATCH", "VOTE", "SHFL",
            "REDUX", "BAR.SYNC", "MEMBAR", "CCTL", "ERRBAR", "STS.128", "LDGDEPBAR", "FENCE"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, cnt = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        cnt[cur] = collections.Counter()
        continue
    mm = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if cur and mm:
        op = mm.group(1)
        cnt[cur]["_total"] += 1
        for p in PREFIXES:
            if op == p or op.startswith(p + ".") or (p.count(".") and op.startswith(p)):
                cnt[cur][p] += 1
arch = re.findall(r"arch = (sm_\w+)", subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout)
print("# %s: %d kernels, cubins %s" % (os.path.basename(lib), len(cnt), sorted(set(arch)) or "sm_100a (see -gencode in the Makefile)"))
for f, c in cnt.items():
    name = subprocess.run(["c++filt", f], capture_output=True, text=True).stdout.strip()
    name = re.sub(r"\(.*", "", name)
    print("%-58s %6d instr  %s" % (name[:58], c["_total"], " ".join("%s=%d" % (k, v) for k, v in sorted(c.items()) if k != "_total")))
