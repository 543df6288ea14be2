"""Small workload for compute-sanitizer runs (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kxpu_b200 as K
from kxpu_b200 import workloads as W
kx = K.Kxpu(0)
text = W.load_pci_ids()
copies = int(os.environ.get("COPIES", "3"))
tab = kx.pciids_load(text * copies)
keys, offs, rows = kx.table_export(tab)
r = kx.lookup(tab, W.cfg2_queries(keys))
names, _, _ = kx.names(tab, r)
print("rows", tab.rows, "hits", int((r >= 0).sum()))
recs = W.cfg3_records(keys, n=20000)
res = kx.classify(recs)
devs = W.cfg5_devices(3000)
print("classify", res["n_accepted"], "json", len(kx.cdi_emit(1, devs)), "yaml", len(kx.cdi_emit(0, devs)))
print("alloc", len(kx.alloc_names(devs["index"])[0]), "lw", len(kx.lw_encode(res["group_ids"][:100])))
tab.free()
