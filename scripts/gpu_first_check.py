"""First GPU check of the pci.ids path: cfg2 parse + lookups + names vs the oracle."""
import hashlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kxpu_b200 as K
from kxpu_b200 import workloads as W
from oracle import oracle as O

text = W.load_pci_ids()
assert hashlib.sha256(text).hexdigest() == W.PCI_IDS_SHA256
kx = K.Kxpu(0)
t0 = time.time(); tab = kx.pciids_load(text); t1 = time.time()
print("rows", tab.rows, "load s", t1 - t0, "timings ms", kx.timings())
keys, offs, rows = kx.table_export(tab)
orows = O.table_build(text)
print("oracle rows", len(orows))
print("keys equal", np.array_equal(keys, orows["key"]), "offs equal", np.array_equal(offs, orows["line_off"]))
names, blob, no = kx.names(tab, rows)
dump = b"".join(b"%04x:%04x\t%s\n" % (k >> 16, k & 0xffff, nm) for k, nm in zip(keys, names))
gold = json.load(open(os.path.join(os.path.dirname(W.PCI_IDS_GZ), "golden.json")))
print("dump sha ok", hashlib.sha256(dump).hexdigest() == gold["dump_sha256"], len(dump))
q = W.cfg2_queries(keys)
r = kx.lookup(tab, q)
ooffs, onames = O.lookup_many(text, q)
line_of_row = dict(zip(rows.tolist(), offs.tolist()))
got = np.array([line_of_row[x] if x >= 0 else -1 for x in r.tolist()], dtype=np.int64)
print("cfg2 lookups equal", np.array_equal(got, ooffs), "hits", int((r >= 0).sum()))
gn, _, _ = kx.names(tab, r)
print("cfg2 names equal", all((a or b"") == b for a, b in zip(onames, gn)))
# x1000 device-resident
n = len(text); copies = int(os.environ.get("COPIES", "1000"))
d_one = kx.dev_alloc(n); kx.upload(d_one, np.frombuffer(text, np.uint8))
d_big = kx.dev_alloc(n * copies); kx.replicate(d_big, d_one, n, copies)
for it in range(5):
    t0 = time.time(); tb = kx.pciids_load_device(d_big, n * copies); t1 = time.time()
    tm = kx.timings()
    print("x%d rows %d wall %.3f ms parse %.3f ms (%.1f GB/s) finalize %.3f ms" % (copies, tb.rows, (t1 - t0) * 1e3, tm[0], n * copies / tm[0] / 1e6, tm[1]))
    k2, o2, r2 = kx.table_export(tb)
    if it == 0:
        print("x1000 table equal", np.array_equal(k2, keys), np.array_equal(o2, offs))
    tb.free()
print("launches", kx.launch_count())
