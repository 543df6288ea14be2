"""Fake /sys/bus/pci/devices tree: every <bdf> entry is a SYMLINK to a
directory holding `vendor`, `device` and the links `driver`, `iommu_group`, like real sysfs."""
import ctypes as C
import json
import os

import numpy as np

from conftest import ROOT

HOST_LIB = os.path.join(ROOT, "kata-xpu-device-plugin_b200", "lib", "libkxpu_host.so")


def make_tree(root, devices):
    """devices: list of dicts(bdf, vendor=b'0x10de\\n'|None, device=..|None, driver='vfio-pci'|None,
    group=214|None, kind='link'|'dir'|'file')."""
    base = os.path.join(root, "bus", "pci", "devices")
    real = os.path.join(root, "devices")
    os.makedirs(base)
    os.makedirs(os.path.join(root, "drivers"))
    os.makedirs(os.path.join(root, "iommu_groups"))
    for d in devices:
        kind = d.get("kind", "link")
        if kind == "file":
            open(os.path.join(base, d["bdf"]), "w").write("x")
            continue
        target = os.path.join(base, d["bdf"]) if kind == "dir" else os.path.join(real, d["bdf"])
        os.makedirs(target)
        if d.get("vendor") is not None:
            open(os.path.join(target, "vendor"), "wb").write(d["vendor"])
        if d.get("device") is not None:
            open(os.path.join(target, "device"), "wb").write(d["device"])
        if d.get("driver") is not None:
            drv = os.path.join(root, "drivers", d["driver"])
            os.makedirs(drv, exist_ok=True)
            os.symlink(drv, os.path.join(target, "driver"))
        if d.get("group") is not None:
            grp = os.path.join(root, "iommu_groups", str(d["group"]))
            os.makedirs(grp, exist_ok=True)
            os.symlink(grp, os.path.join(target, "iommu_group"))
        if kind == "link":
            os.symlink(target, os.path.join(base, d["bdf"]))
    return base


def host_lib():
    L = C.CDLL(HOST_LIB)
    L.kxh_gather.restype = C.c_int
    L.kxh_gather.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_char_p, C.c_size_t]
    L.kxh_gather_fast.restype = C.c_int
    L.kxh_gather_fast.argtypes = [C.c_char_p, C.c_uint, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_char_p, C.c_size_t]
    L.kxh_new.restype = C.c_void_p
    L.kxh_new.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p]
    L.kxh_free.argtypes = [C.c_void_p]
    L.kxh_init.restype = C.c_int
    L.kxh_init.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_size_t]
    L.kxh_allocate.restype = C.c_int
    L.kxh_allocate.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_size_t]
    L.kxh_add_plugin.restype = C.c_int
    L.kxh_add_plugin.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p]
    L.kxh_health_start.restype = C.c_void_p
    L.kxh_health_start.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_size_t]
    L.kxh_health_poll.restype = C.c_int
    L.kxh_health_poll.argtypes = [C.c_void_p, C.c_int]
    L.kxh_health_stop.argtypes = [C.c_void_p]
    L.kxh_devs.restype = C.c_int
    L.kxh_devs.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_size_t]
    L.kxh_list_and_watch.restype = C.c_int
    L.kxh_list_and_watch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
    return L


def gather(base, dtype, cap=4096):
    L = host_lib()
    recs = np.zeros(cap, dtype=dtype)
    n = C.c_size_t(0)
    err = C.create_string_buffer(512)
    rc = L.kxh_gather(base.encode(), recs.ctypes.data, cap, C.byref(n), err, 512)
    if rc != 0:
        raise RuntimeError(err.value.decode())
    return recs[:n.value]


def gather_fast(base, dtype, threads=0, cap=4096):
    """Same records through the batched / threaded gather."""
    L = host_lib()
    recs = np.zeros(cap, dtype=dtype)
    n = C.c_size_t(0)
    err = C.create_string_buffer(512)
    rc = L.kxh_gather_fast(base.encode(), threads, recs.ctypes.data, cap, C.byref(n), err, 512)
    if rc != 0:
        raise RuntimeError(err.value.decode())
    return recs[:n.value]


class HostPlugin:
    def __init__(self, kx, base, pciids, cdi_dir):
        self.L = host_lib()
        self.h = self.L.kxh_new(kx.ctx, base.encode(), pciids.encode(), cdi_dir.encode())

    def init(self, fmt="YAML"):
        buf = C.create_string_buffer(1 << 22)
        rc = self.L.kxh_init(self.h, fmt.encode(), buf, len(buf))
        if rc < 0:
            raise RuntimeError(buf.value.decode())
        return json.loads(buf.value.decode())

    def allocate(self, ids):
        buf = C.create_string_buffer(1 << 16)
        rc = self.L.kxh_allocate(self.h, ",".join(ids).encode(), buf, len(buf))
        if rc < 0:
            raise RuntimeError(buf.value.decode())
        return json.loads(buf.value.decode())

    def list_and_watch(self, idx):
        out = np.empty(1 << 20, np.uint8)
        rc = self.L.kxh_list_and_watch(self.h, idx, out.ctypes.data, out.size)
        assert rc >= 0
        return out[:rc].tobytes()

    def close(self):
        self.L.kxh_free(self.h)
