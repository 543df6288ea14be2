"""Device time of the parse / resolve / finalize kernels against the shard size (pci.ids x COPIES on one
GPU): the fixed cost that limits strong scaling.  KXPU_RCH forces the chunks per range."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kxpu_b200 as K
from kxpu_b200 import workloads as W

B = K.binding
kx = K.Kxpu(0)
text = np.frombuffer(W.load_pci_ids(), np.uint8)
n1 = len(text)
d_one = kx.dev_alloc(n1)
kx.upload(d_one, text)
for copies in [int(c) for c in os.environ.get("COPIES", "1000,500,250,125,60,16,4,1").split(",")]:
    d = kx.dev_alloc(n1 * copies)
    kx.replicate(d, d_one, n1, copies)
    ts = []
    for it in range(12):
        t = kx.pciids_load_device(d, n1 * copies)
        ts.append(kx.timings())
        t.free()
    tm = np.median(np.array(ts[4:]), axis=0)
    n = n1 * copies
    print("x%-5d %6.1f MB  parse %7.1f us (%6.0f GB/s)  resolve %6.1f us  finalize %6.1f us" %
          (copies, n / 1e6, tm[B.T_PARSE] * 1e3, n / tm[B.T_PARSE] / 1e6, tm[B.T_RESOLVE] * 1e3, tm[B.T_FINALIZE] * 1e3), flush=True)
    kx.dev_free(d)
