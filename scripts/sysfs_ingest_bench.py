"""SURVEY 8(f) row 2, measurement: sysfs ingestion on a synthetic /sys/bus/pci/devices tree --
the reference-shaped walk (lstat + fopen/read/close x2 + readlink x2 per function, absolute paths)
against the batched gather (getdents64 d_type, openat/readlinkat relative to one dirfd, threads).
CPU only; both produce the identical record table (tests/test_host.py)."""
import os, sys, tempfile, time, shutil
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
import numpy as np
import fake_sysfs
from test_host import _synthetic_devices
from kxpu_b200.binding import DEVREC_DTYPE

n = int(os.environ.get("DEVICES", "20000"))
tmp = tempfile.mkdtemp(prefix="kxpu_sysfs_")
try:
    base = fake_sysfs.make_tree(tmp, _synthetic_devices(n, 1)[:n])
    def best(fn, *a, **kw):
        ts = []
        for _ in range(5):
            t0 = time.time(); r = fn(base, DEVREC_DTYPE, *a, cap=n + 64, **kw); ts.append(time.time() - t0)
        return min(ts), r
    t_walk, ref = best(fake_sysfs.gather)
    print("devices %d  walk (reference-shaped): %.1f ms  %.0f records/s" % (n, t_walk * 1e3, len(ref) / t_walk))
    for th in (1, 2, 4, 8):
        t, r = best(fake_sysfs.gather_fast, th)
        assert r.tobytes() == ref.tobytes()
        print("  batched gather, %d thread(s): %.1f ms  %.0f records/s  x%.2f" % (th, t * 1e3, len(r) / t, t_walk / t))
finally:
    shutil.rmtree(tmp, ignore_errors=True)
