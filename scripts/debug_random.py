import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kxpu_b200 as K
from oracle import oracle as O
kx = K.Kxpu(0)
rng = np.random.default_rng(2024)
for trial in range(12):
    lines = []
    for _ in range(int(rng.integers(5, 400))):
        r = rng.random()
        v = int(rng.integers(0, 40)) * 0x0101
        if r < 0.2: lines.append(b"%04x  Vendor %d\n" % (v, v))
        elif r < 0.7: lines.append(b"\t%04x  Dev.%d / x\n" % (int(rng.integers(0, 60)), int(rng.integers(0, 1000))))
        elif r < 0.8: lines.append(b"\t\t%04x %04x  Sub\n" % (v, v))
        elif r < 0.85: lines.append(b"# c\n")
        elif r < 0.88: lines.append(b"\n")
        elif r < 0.91: lines.append(b"C %02x  Class\n" % int(rng.integers(0, 255)))
        elif r < 0.94: lines.append(b"\t%04X  UPPER\n" % int(rng.integers(0xa000, 0xffff)))
        elif r < 0.97: lines.append(b"\t12\n")
        else: lines.append(b"%04x\n" % v)
    text = b"".join(lines)
    if trial % 3 == 0: text = text.rstrip(b"\n")
    keys = [(int(rng.integers(0, 40)) * 0x0101 << 16) | int(rng.integers(0, 60)) for _ in range(40)]
    tab = kx.pciids_load(text)
    k, o, r = kx.table_export(tab)
    orow = O.table_build(text)
    if not (np.array_equal(k, orow["key"]) and np.array_equal(o, orow["line_off"])):
        gs = set(zip(k.tolist(), o.tolist())); os_ = set(zip(orow["key"].tolist(), orow["line_off"].tolist()))
        print("trial", trial, "len", len(text), "gpu-only", [(hex(a), b) for a, b in sorted(gs - os_, key=lambda x: x[1])][:5],
              "oracle-only", [(hex(a), b) for a, b in sorted(os_ - gs, key=lambda x: x[1])][:5])
        for a, b in sorted(gs - os_, key=lambda x: x[1])[:2]:
            print(repr(text[max(0, b - 60):b + 40]))
    tab.free()
