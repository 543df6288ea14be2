#!/usr/bin/env python3
"""Regenerates tests/golden/* in the BUILD container (needs /root/reference).

  pci.ids.gz        the reference's bundled PCI ID database (utils/pci.ids, v2024.06.23;
                    public data, GPL-2+/BSD-3 per its own header), gzip'ed input fixture for
                    BASELINE.json configs[1] and configs[3].  sha256 of the plain text is
                    pinned in golden.json.
  golden.json       self-pinned hashes of the oracle's canonical dump and
                    spot values.
  cfg1.yaml/.json   the one-device CDI documents of SURVEY.md 8a-fmt.

The reference has no tests/golden vectors of its own and cannot be executed here (Go),
so these pin the ORACLE across sessions ("parity unpinned", see oracle/kxpu_oracle.c).
"""
import gzip
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import oracle as O  # noqa: E402

SRC = "/root/reference/utils/pci.ids"


def main():
    text = open(SRC, "rb").read()
    with open(os.path.join(HERE, "pci.ids.gz"), "wb") as f:
        with gzip.GzipFile(fileobj=f, mode="wb", mtime=0, compresslevel=9) as g:
            g.write(text)
    rows = O.table_build(text)
    dump = b"".join(b"%04x:%04x\t%s\n" % (r["key"] >> 16, r["key"] & 0xFFFF, O.row_name(text, int(r["line_off"])))
                    for r in rows)
    nv = b"".join(l + b"\n" for l in dump.split(b"\n") if l.startswith(b"10de:"))
    spots = {}
    for k in [0x10de2330, 0x10de2331, 0x10de20b0, 0x10de20b5, 0x10de2684, 0x10de1db4, 0x10de0020, 0x10de28e0,
              0x10de0fb9, 0x1d0fefa1, 0x80861572, 0x100273bf, 0x103cb204, 0x10de2901, 0xffff0000, 0x11ab2b42]:
        off, nm = O.device_name(text, k)
        spots["%08x" % k] = {"line_off": off, "name": None if nm is None else nm.decode()}
    gold = {
        "pci_ids_sha256": hashlib.sha256(text).hexdigest(),
        "pci_ids_bytes": len(text),
        "rows": int(len(rows)),
        "dump_bytes": len(dump),
        "dump_sha256": hashlib.sha256(dump).hexdigest(),
        "nvidia_rows": nv.count(b"\n"),
        "nvidia_dump_sha256": hashlib.sha256(nv).hexdigest(),
        "spots": spots,
    }
    json.dump(gold, open(os.path.join(HERE, "golden.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(gold, indent=1)[:600])


if __name__ == "__main__":
    main()
