"""Full path of the parse (every vendor block a first occurrence, nothing can be skipped): synthetic
pci.ids-shaped text without replication (workloads.synthetic_pci_ids) -- parity against the oracle,
device time per stage.  VENDORS / DEVS pick the size (65536 x 190 ~ 0.96 GB, 12.4 M keys)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kxpu_b200 as K
from kxpu_b200 import workloads as W
from oracle import oracle as O

B = K.binding
kx = K.Kxpu(0)
for nv, nd in [(int(x) for x in s.split("x")) for s in os.environ.get("SIZES", "2388x8,40000x24,65536x190").split(",")]:
    text = W.synthetic_pci_ids(nv, nd)
    n = len(text)
    d = kx.dev_alloc(n)
    kx.upload(d, text)
    ITERS = int(os.environ.get("ITERS", "5"))
    for it in range(ITERS):
        tb = kx.pciids_load_device(d, n)
        tm = kx.timings()
        print("%dx%d iter %d: %d B, %d rows | parse %.3f ms (%.1f GB/s) resolve %.3f ms finalize %.3f ms" %
              (nv, nd, it, n, tb.rows, tm[B.T_PARSE], n / tm[B.T_PARSE] / 1e6, tm[B.T_RESOLVE], tm[B.T_FINALIZE]), flush=True)
        if it < ITERS - 1:
            tb.free()
    if os.environ.get("NO_ORACLE"):
        tb.free()
        kx.dev_free(d)
        continue
    keys, offs, rows = kx.table_export(tb)
    t0 = time.time()
    orows = O.table_build(text)
    same = np.array_equal(keys, orows["key"]) and np.array_equal(offs, orows["line_off"])
    names, _, _ = kx.names(tb, rows[:500])
    nm_ok = all(nm == O.row_name(text, int(o)) for nm, o in zip(names, offs[:500]))
    print("   oracle %.2f s; table equal %s, names equal %s" % (time.time() - t0, same, nm_ok), flush=True)
    tb.free()
    kx.dev_free(d)
