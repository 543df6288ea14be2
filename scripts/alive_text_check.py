"""Worst case for the alive-first parse: a text in which EVERY vendor block is a first occurrence
(nothing can be skipped) -- parity against the oracle and device time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kxpu_b200 as K
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import oracle as O

nv = int(os.environ.get("VENDORS", "40000"))
nd = int(os.environ.get("DEVS", "24"))
rng = np.random.default_rng(1)
parts = []
for v in range(nv):
    parts.append(b"%04x  Vendor number %d\n" % (v, v))
    ds = rng.choice(65536, nd, replace=False)
    parts.append(b"".join(b"\t%04x  Device %d\n\t\t%04x %04x  Sub\n" % (d, d, v, d) for d in ds))
text = b"".join(parts)
n = len(text)
kx = K.Kxpu(0)
d = kx.dev_alloc(n)
kx.upload(d, np.frombuffer(text, np.uint8))
for it in range(4):
    tb = kx.pciids_load_device(d, n)
    tm = kx.timings()
    print("iter %d bytes %d rows %d parse %.3f ms (%.1f GB/s) resolve %.3f ms finalize %.3f ms" % (it, n, tb.rows, tm[0], n / tm[0] / 1e6, tm[7], tm[1]), flush=True)
    if it < 3:
        tb.free()
keys, offs, rows = kx.table_export(tb)
t0 = time.time()
orows = O.table_build(text)
print("oracle %.2f s; keys equal %s offsets equal %s" % (time.time() - t0, np.array_equal(keys, orows["key"]), np.array_equal(offs, orows["line_off"])))
