"""The other rows of the hot path for an ncu launch list: cfg3 classify, cfg5 CDI emit (JSON / YAML),
cfg2 (real pci.ids once + 1024 lookups), Allocate names, ListAndWatch bytes -- a few warm rounds each."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kxpu_b200 as K
from kxpu_b200 import workloads as W

B = K.binding
kx = K.Kxpu(0)
text = W.load_pci_ids()
t = kx.pciids_load(text)
present, _, _ = kx.table_export(t)
t.free()
recs = W.cfg3_records(present)
devs = W.cfg5_devices()
q2 = W.cfg2_queries(present)
for rnd in range(int(os.environ.get("ROUNDS", "3"))):
    res = kx.classify(recs)
    j = kx.cdi_emit(B.FMT_JSON, devs)
    y = kx.cdi_emit(B.FMT_YAML, devs)
    t = kx.pciids_load(np.frombuffer(text, np.uint8))
    r = kx.lookup(t, q2)
    nm, _, _ = kx.names(t, r)
    t.free()
    an = kx.alloc_names(devs["index"])
    lw = kx.lw_encode(res["group_ids"][:100000])
    tm = kx.timings()
print("ok", res["n_accepted"], len(j), len(y), int((r >= 0).sum()), len(an[0]), len(lw))
