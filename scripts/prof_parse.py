"""Driver for ncu captures of the parse kernel: x COPIES pci.ids resident in HBM, ITERS loads."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kxpu_b200 as K
from kxpu_b200 import workloads as W

copies = int(os.environ.get("COPIES", "1000"))
iters = int(os.environ.get("ITERS", "4"))
text = W.load_pci_ids()
n = len(text)
kx = K.Kxpu(0)
d_one = kx.dev_alloc(n)
kx.upload(d_one, np.frombuffer(text, np.uint8))
d_big = kx.dev_alloc(n * copies)
kx.replicate(d_big, d_one, n, copies)
for it in range(iters):
    tb = kx.pciids_load_device(d_big, n * copies)
    tm = kx.timings()
    print("iter %d rows %d parse %.3f ms %.1f GB/s resolve %.3f ms finalize %.3f ms" % (it, tb.rows, tm[0], n * copies / tm[0] / 1e6, tm[7], tm[1]))
    tb.free()
