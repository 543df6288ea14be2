"""cfg2 (the real pci.ids once + 1024 keys) through kxpu_pciids_join with KXPU_TRACE_SMALL=1: SM cycles of every
phase of the cooperative small-text kernel (stderr), a few warm calls."""
import os, sys
os.environ["KXPU_TRACE_SMALL"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kxpu_b200 as K
from kxpu_b200 import workloads as W

kx = K.Kxpu(0)
text = np.frombuffer(W.load_pci_ids(), np.uint8)
t = kx.pciids_load(text)
present, _, _ = kx.table_export(t)
t.free()
q2 = W.cfg2_queries(present)
for it in range(6):
    t, rows = kx.pciids_join(text, q2)
    t.free()
print("hits", int((rows >= 0).sum()))
