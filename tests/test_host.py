"""Host-side mirror of pkg/device_plugin (C++, kata-xpu-device-plugin_b200/host): the sysfs walk on
CPU, and the whole InitiateDevicePlugin / Allocate flow through the GPU path on a fake sysfs."""
import os

import numpy as np
import pytest

import fake_sysfs
from conftest import GOLDEN

DEVICES = [
    dict(bdf="0000:c5:00.0", vendor=b"0x10de\n", device=b"0x2330\n", driver="vfio-pci", group=215),
    dict(bdf="0000:c1:00.0", vendor=b"0x10de\n", device=b"0x2330\n", driver="vfio-pci", group=214),
    dict(bdf="0000:3d:00.0", vendor=b"0x10de\n", device=b"0x20b5\n", driver="vfio-pci", group=75),
    dict(bdf="0000:3d:00.1", vendor=b"0x10de\n", device=b"0x1aef\n", driver="vfio-pci", group=75),  # same group: audio fn
    dict(bdf="0000:41:00.0", vendor=b"0x10de\n", device=b"0x2901\n", driver="vfio-pci", group=76),   # not in pci.ids
    dict(bdf="0000:00:1f.0", vendor=b"0x8086\n", device=b"0x1572\n", driver="ixgbe", group=3),
    dict(bdf="0000:81:00.0", vendor=b"0x10de\n", device=b"0x2684\n", driver="nvidia", group=90),     # wrong driver
    dict(bdf="0000:82:00.0", vendor=b"0x10de\n", device=b"0x2684\n", driver=None, group=91),         # unbound
    dict(bdf="0000:83:00.0", vendor=None, device=None, driver=None, group=None),                     # unreadable
    dict(bdf="zz_a_directory", kind="dir", vendor=b"0x10de\n", device=b"0x2330\n", driver="vfio-pci", group=7),
]


def test_gather_walk_order_and_raw_bytes(tmp_path, workloads):
    """CPU: createIommuDeviceMap's walk (lexical order, symlink entries are devices, real
    directories are descended into) and the raw bytes it hands to kxpu_classify."""
    from kxpu_b200.binding import DEVREC_DTYPE
    base = fake_sysfs.make_tree(str(tmp_path), DEVICES)
    recs = fake_sysfs.gather(base, DEVREC_DTYPE)
    names = [r["bdf"] for r in recs]
    # entries in lexical order; the directory is descended and its inner names become records
    assert names[:9] == sorted(d["bdf"].encode() for d in DEVICES if d.get("kind") != "dir")
    assert names[9:] == [b"device", b"driver", b"iommu_group", b"vendor"]
    assert all(r["flags"] & 1 for r in recs[9:])  # basePath/<inner name>/vendor does not exist
    r = recs[names.index(b"0000:c1:00.0")]
    assert bytes(r["vendor_txt"][:7]) == b"0x10de\n" and r["vendor_len"] == 7
    assert bytes(r["device_txt"][:7]) == b"0x2330\n" and r["driver"] == b"vfio-pci" and r["iommu_group"] == 214 and r["flags"] == 0
    assert recs[names.index(b"0000:82:00.0")]["flags"] == 2
    assert recs[names.index(b"0000:83:00.0")]["flags"] & 1


def test_gather_matches_oracle_classification(tmp_path, oracle):
    from kxpu_b200.binding import DEVREC_DTYPE
    base = fake_sysfs.make_tree(str(tmp_path), DEVICES)
    recs = fake_sysfs.gather(base, DEVREC_DTYPE)
    res = oracle.classify(recs)
    acc = {r["bdf"]: int(a) for r, a in zip(recs, res["accept_index"]) if a != 0xFFFFFFFF}
    assert acc == {b"0000:3d:00.0": 0, b"0000:3d:00.1": 1, b"0000:41:00.0": 2, b"0000:c1:00.0": 3, b"0000:c5:00.0": 4}
    assert list(res["group_ids"]) == [75, 76, 214, 215]


@pytest.mark.gpu
def test_initiate_device_plugin_flow(tmp_path, kx, workloads, pci_text):
    """cfg1-style plumbing on a fake sysfs with several devices, through the GPU path."""
    base = fake_sysfs.make_tree(str(tmp_path), DEVICES)
    pciids = tmp_path / "pci.ids"
    pciids.write_bytes(pci_text)
    cdi = tmp_path / "cdi"
    cdi.mkdir()
    hp = fake_sysfs.HostPlugin(kx, base, str(pciids), str(cdi) + "/")
    st = hp.init("YAML")
    assert st["iommuMap"] == [["75", [["0000:3d:00.0", 0], ["0000:3d:00.1", 1]]], ["76", [["0000:41:00.0", 2]]],
                              ["214", [["0000:c1:00.0", 3]]], ["215", [["0000:c5:00.0", 4]]]]
    # a group belongs to the device id of its FIRST member (device_plugin.go:162-170)
    assert st["deviceMap"] == [["20b5", ["75"]], ["2901", ["76"]], ["2330", ["214", "215"]]]
    plugins = {p["name"]: p for p in st["plugins"]}
    assert set(plugins) == {"GA100_A100_PCIE_80GB", "2901", "GH100_H100_SXM5_80GB"}  # unknown id falls back to the raw id
    h100 = plugins["GH100_H100_SXM5_80GB"]
    assert h100["resource"] == "nvidia.com/GH100_H100_SXM5_80GB"
    assert h100["socket"] == "/var/lib/kubelet/device-plugins/kata-xpu-GH100_H100_SXM5_80GB.sock"
    assert h100["devs"] == [["214", "Healthy"], ["215", "Healthy"]]
    y = open(st["cdiFile"], "rb").read()
    assert st["cdiFile"].endswith("cdi-vfio-xxxx.yaml")
    assert b'bdf: "0000:41:00.0"' in y and b"bdf: 0000:c1:00.0" in y and b"cdi.k8s.io/vfio75: nvidia.com/gpu=1" in y
    import yaml
    doc = yaml.safe_load(y)
    assert [d["name"] for d in doc["devices"]] == ["0", "1", "2", "3", "4"]
    # Allocate: both functions of group 75, then 214
    r = hp.allocate(["75", "214"])
    assert r == {"envs": {"KUBERNETES_CDI_VENDOR_CLASS": "nvidia.com/gpu"},
                 "cdi_devices": ["nvidia.com/gpu=0", "nvidia.com/gpu=1", "nvidia.com/gpu=3"]}
    assert hp.allocate(["9999"]) == {"envs": {"KUBERNETES_CDI_VENDOR_CLASS": "nvidia.com/gpu"}, "cdi_devices": []}
    idx = [p["name"] for p in st["plugins"]].index("GH100_H100_SXM5_80GB")
    assert hp.list_and_watch(idx) == b"\x0a\x0e\x0a\x03214\x12\x07Healthy\x0a\x0e\x0a\x03215\x12\x07Healthy"
    # re-validation: the device moved to another IOMMU group -> error naming the bdf
    link = os.path.join(base, "0000:c1:00.0", "iommu_group")
    os.unlink(link)
    os.symlink(os.path.join(str(tmp_path), "iommu_groups", "215"), link)
    with pytest.raises(RuntimeError, match="invalid allocation request: unknown device: 0000:c1:00.0"):
        hp.allocate(["214"])
    hp.close()


@pytest.mark.gpu
def test_cfg1_single_gpu_golden_files(tmp_path, kx, pci_text):
    """BASELINE.json configs[0] exactly as SURVEY.md 8(d) describes it."""
    base = fake_sysfs.make_tree(str(tmp_path), [dict(bdf="0000:c1:00.0", vendor=b"0x10de\n", device=b"0x2330\n",
                                                     driver="vfio-pci", group=214)])
    pciids = tmp_path / "pci.ids"
    pciids.write_bytes(pci_text)
    cdi = tmp_path / "cdi"
    cdi.mkdir()
    hp = fake_sysfs.HostPlugin(kx, base, str(pciids), str(cdi) + "/")
    st = hp.init("YAML")
    assert st["iommuMap"] == [["214", [["0000:c1:00.0", 0]]]] and st["deviceMap"] == [["2330", ["214"]]]
    assert st["plugins"][0]["resource"] == "nvidia.com/GH100_H100_SXM5_80GB"
    assert open(st["cdiFile"], "rb").read() == open(os.path.join(GOLDEN, "cfg1.yaml"), "rb").read()
    st = hp.init("JSON")
    assert open(st["cdiFile"], "rb").read() == open(os.path.join(GOLDEN, "cfg1.json"), "rb").read()
    assert hp.allocate(["214"])["cdi_devices"] == ["nvidia.com/gpu=0"]
    hp.close()


# ---------------------------------------------------------------------------------------------
# health events -> batched re-emit of the ListAndWatch list
# (generic_device_plugin.go:389-457 healthCheck, :222-250 ListAndWatch)
# ---------------------------------------------------------------------------------------------
def _devs(L, h, idx):
    import ctypes as C
    buf = C.create_string_buffer(1 << 16)
    assert L.kxh_devs(h, idx, buf, len(buf)) >= 0
    return dict(kv.split("=") for kv in buf.value.decode().split(",") if kv)


def test_health_watcher_marks_removed_and_renamed_groups_unhealthy(tmp_path):
    """CPU: inotify on devicePath/<group>; Remove and Rename mark the device Unhealthy, a whole
    burst of events is one batch; without the directory watch a re-Create is not seen (exactly
    like the reference, which only watches the device paths and the socket directory)."""
    import ctypes as C
    L = fake_sysfs.host_lib()
    vfio = tmp_path / "vfio"
    vfio.mkdir()
    ids = ["214", "215", "75", "76"]
    for i in ids:
        (vfio / i).write_text("")
    h = L.kxh_new(None, b"/nonexistent", b"/nonexistent", b"/nonexistent")
    idx = L.kxh_add_plugin(h, b"GH100_H100_SXM5_80GB", (str(vfio) + "/").encode(), ",".join(ids).encode())
    err = C.create_string_buffer(512)
    w = L.kxh_health_start(h, idx, 0, err, len(err))
    assert w, err.value
    try:
        assert L.kxh_health_poll(w, 0) == 0 and set(_devs(L, h, idx).values()) == {"Healthy"}
        os.remove(vfio / "214")
        os.rename(vfio / "75", vfio / "75.moved")
        os.remove(vfio / "76")
        assert L.kxh_health_poll(w, 1000) == 3  # one batch
        assert _devs(L, h, idx) == {"214": "Unhealthy", "215": "Healthy", "75": "Unhealthy", "76": "Unhealthy"}
        assert L.kxh_health_poll(w, 0) == 0
        (vfio / "214").write_text("")  # re-created: not seen without a directory watch
        assert L.kxh_health_poll(w, 50) == 0 and _devs(L, h, idx)["214"] == "Unhealthy"
    finally:
        L.kxh_health_stop(w)
        L.kxh_free(h)
    # a device path that does not exist makes healthCheck fail (:426-429)
    h = L.kxh_new(None, b"/nonexistent", b"/nonexistent", b"/nonexistent")
    idx = L.kxh_add_plugin(h, b"X", (str(vfio) + "/").encode(), b"999")
    assert not L.kxh_health_start(h, idx, 0, err, len(err)) and b"999" in err.value
    L.kxh_free(h)


def test_health_watcher_with_directory_watch_sees_creates(tmp_path):
    import ctypes as C
    L = fake_sysfs.host_lib()
    vfio = tmp_path / "vfio"
    vfio.mkdir()
    for i in ("1", "2"):
        (vfio / i).write_text("")
    h = L.kxh_new(None, b"/nonexistent", b"/nonexistent", b"/nonexistent")
    idx = L.kxh_add_plugin(h, b"X", str(vfio).encode(), b"1,2")
    err = C.create_string_buffer(512)
    w = L.kxh_health_start(h, idx, 1, err, len(err))
    assert w, err.value
    try:
        os.remove(vfio / "1")
        assert L.kxh_health_poll(w, 1000) == 1 and _devs(L, h, idx) == {"1": "Unhealthy", "2": "Healthy"}
        (vfio / "1").write_text("")
        (vfio / "unrelated").write_text("")
        assert L.kxh_health_poll(w, 1000) == 1 and _devs(L, h, idx) == {"1": "Healthy", "2": "Healthy"}
        os.remove(vfio / "1")  # the re-created file is watched again
        assert L.kxh_health_poll(w, 1000) == 1 and _devs(L, h, idx)["1"] == "Unhealthy"
        # flapping inside one batch: only the net change counts towards a re-emit
        (vfio / "1").write_text("")
        os.remove(vfio / "1")
        (vfio / "1").write_text("")
        L.kxh_health_poll(w, 1000)
        assert _devs(L, h, idx)["1"] == "Healthy"
    finally:
        L.kxh_health_stop(w)
        L.kxh_free(h)


@pytest.mark.gpu
def test_health_batch_reemits_list_and_watch_bytes(tmp_path, kx, oracle):
    """GPU: after a batch of health events the re-encoded ListAndWatchResponse equals the oracle's
    encoding of the same list with the same health flags (what the reference's last s.Send carries)."""
    import ctypes as C
    L = fake_sysfs.host_lib()
    vfio = tmp_path / "vfio"
    vfio.mkdir()
    ids = [str(1000 + i) for i in range(300)]
    for i in ids:
        (vfio / i).write_text("")
    h = L.kxh_new(kx.ctx, b"/nonexistent", b"/nonexistent", b"/nonexistent")
    idx = L.kxh_add_plugin(h, b"X", str(vfio).encode(), ",".join(ids).encode())
    err = C.create_string_buffer(512)
    w = L.kxh_health_start(h, idx, 0, err, len(err))
    assert w, err.value
    try:
        out = (C.c_uint8 * (1 << 16))()
        n = L.kxh_list_and_watch(h, idx, out, len(out))
        g = np.array([int(i) for i in ids], np.uint32)
        assert bytes(out[:n]) == oracle.lw_encode(g)
        gone = ids[3::7]
        for i in gone:
            os.remove(vfio / i)
        assert L.kxh_health_poll(w, 1000) == len(gone)
        n = L.kxh_list_and_watch(h, idx, out, len(out))
        healthy = np.array([0 if i in set(gone) else 1 for i in ids], np.uint8)
        assert bytes(out[:n]) == oracle.lw_encode(g, healthy)
    finally:
        L.kxh_health_stop(w)
        L.kxh_free(h)


# ---------------------------------------------------------------------------------------------
# batched sysfs ingestion == the reference-shaped walk, record for record
# ---------------------------------------------------------------------------------------------
def _synthetic_devices(n, seed):
    rng = np.random.default_rng(seed)
    devs = []
    for i in range(n):
        bdf = "%04x:%02x:%02x.%x" % (i >> 16, (i >> 8) & 0xff, (i >> 3) & 0x1f, i & 7)
        r = rng.random()
        d = dict(bdf=bdf, vendor=b"0x10de\n" if r < 0.6 else b"0x8086\n", device=b"0x%04x\n" % int(rng.integers(0, 0x3000)),
                 driver="vfio-pci" if rng.random() < 0.8 else ("nvidia" if rng.random() < 0.5 else None), group=i >> 3)
        if r > 0.97:
            d["vendor"] = None            # unreadable vendor file
        elif r > 0.94:
            d["device"] = None
        elif r > 0.92:
            d["group"] = None
        elif r > 0.91:
            d["vendor"] = b"0x1\n"        # short id file
        devs.append(d)
    devs.append(dict(bdf="zz_real_dir", kind="dir", vendor=b"0x10de\n", device=b"0x2330\n", driver="vfio-pci", group=7))
    devs.append(dict(bdf="zz_plain_file", kind="file"))
    return devs


@pytest.mark.parametrize("threads", [1, 4, 0])
def test_fast_gather_equals_walk(tmp_path, threads):
    from kxpu_b200.binding import DEVREC_DTYPE
    base = fake_sysfs.make_tree(str(tmp_path / "a"), DEVICES)
    assert fake_sysfs.gather_fast(base, DEVREC_DTYPE, threads).tobytes() == fake_sysfs.gather(base, DEVREC_DTYPE).tobytes()
    base = fake_sysfs.make_tree(str(tmp_path / "b"), _synthetic_devices(1500, 3))
    slow = fake_sysfs.gather(base, DEVREC_DTYPE, cap=4096)
    fast = fake_sysfs.gather_fast(base, DEVREC_DTYPE, threads, cap=4096)
    assert len(slow) == 1500 + 4 + 1 and fast.tobytes() == slow.tobytes()


def test_fast_gather_error_behaviour_equals_walk(tmp_path):
    """An entry the record cannot carry (non-canonical iommu_group of an NVIDIA vfio-pci function) is
    skipped like a read error -- flag KXPU_REC_IOMMU_ERR -- and never stops the walk; both gathers agree.
    On entries the reference would not accept anyway (other vendor / driver) nothing is read at all."""
    from kxpu_b200.binding import DEVREC_DTYPE, REC_IOMMU_ERR
    devs = _synthetic_devices(300, 5)
    for d in devs[198:203]:
        d.update(vendor=b"0x10de\n", driver="vfio-pci", device=b"0x2330\n", group=90)
    devs[201].update(vendor=b"0x8086\n")
    base = fake_sysfs.make_tree(str(tmp_path), devs)
    for k in (200, 201):  # 200: NVIDIA + vfio-pci -> skipped with a log; 201: other vendor -> the link is never read
        victim = os.path.join(base, devs[k]["bdf"], "iommu_group")
        if os.path.lexists(victim):
            os.remove(victim)
        os.symlink("/somewhere/not-a-number", victim)
    slow = fake_sysfs.gather(base, DEVREC_DTYPE)
    fast = fake_sysfs.gather_fast(base, DEVREC_DTYPE, threads=4)
    assert slow.tobytes() == fast.tobytes() and len(slow) >= 300
    by_bdf = {r["bdf"]: r for r in slow}
    assert by_bdf[devs[200]["bdf"].encode()]["flags"] & REC_IOMMU_ERR
    assert by_bdf[devs[199]["bdf"].encode()]["flags"] == 0 and by_bdf[devs[202]["bdf"].encode()]["flags"] == 0
    assert by_bdf[devs[201]["bdf"].encode()]["flags"] == 0 and by_bdf[devs[201]["bdf"].encode()]["driver"] == b""
    with pytest.raises(RuntimeError):
        fake_sysfs.gather_fast(str(tmp_path / "missing"), DEVREC_DTYPE)


def test_bind_watcher_counts_pci_bind_events():
    """BindWatcher (second half): the generation moves on pci bind / unbind / add / remove
    uevents and on nothing else; the real NETLINK_KOBJECT_UEVENT socket opens where the box allows it."""
    import ctypes as C
    L = fake_sysfs.host_lib()
    L.kxh_uevent_feed.restype = C.c_uint64
    L.kxh_uevent_feed.argtypes = [C.POINTER(C.c_void_p), C.c_char_p, C.c_size_t]
    L.kxh_uevent_free.argtypes = [C.c_void_p]
    w = C.c_void_p()

    def feed(*fields):
        m = b"\0".join(fields) + b"\0"
        return L.kxh_uevent_feed(C.byref(w), m, len(m))
    assert feed(b"add@/devices/virtual/net/x", b"ACTION=add", b"SUBSYSTEM=net") == 0
    assert feed(b"bind@/devices/pci0000:00/0000:00:1f.0", b"ACTION=bind", b"DEVPATH=/devices/pci0000:00/0000:00:1f.0",
                b"SUBSYSTEM=pci", b"DRIVER=vfio-pci") == 1
    assert feed(b"unbind@/devices/x", b"ACTION=unbind", b"SUBSYSTEM=pci") == 2
    assert feed(b"change@/devices/x", b"ACTION=change", b"SUBSYSTEM=pci") == 2
    assert feed(b"remove@/devices/x", b"SUBSYSTEM=pci", b"ACTION=remove") == 3
    assert feed(b"bind@/devices/x", b"ACTION=bind", b"SUBSYSTEM=usb") == 3
    assert L.kxh_uevent_feed(C.byref(w), b"ACTION=bind\0SUBSYSTEM=pci", 26) == 4  # unterminated last field
    L.kxh_uevent_free(w)
    assert L.kxh_uevent_socket_ok() in (0, -1)  # sandboxes may forbid netlink sockets


@pytest.mark.gpu
def test_allocate_snapshot_validation_follows_the_generation(tmp_path, kx, pci_text):
    """snapshotValidation (default off): while the bind generation equals the one recorded at discovery Allocate
    answers from the snapshot -- no sysfs reads --, any change (or an unhealthy watcher) falls back to the
    reference's live reads, which then see the re-bound device (generic_device_plugin.go:329-338)."""
    import ctypes as C
    base = fake_sysfs.make_tree(str(tmp_path), DEVICES)
    pciids = tmp_path / "pci.ids"
    pciids.write_bytes(pci_text)
    cdi = tmp_path / "cdi"
    cdi.mkdir()
    hp = fake_sysfs.HostPlugin(kx, base, str(pciids), str(cdi) + "/")
    gen, healthy = C.c_uint64(7), C.c_int(1)
    hp.L.kxh_snapshot_enable.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_int)]
    hp.L.kxh_validation_counts.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    hp.L.kxh_snapshot_enable(hp.h, C.byref(gen), C.byref(healthy))

    def counts():
        a, b = C.c_uint64(0), C.c_uint64(0)
        hp.L.kxh_validation_counts(hp.h, C.byref(a), C.byref(b))
        return a.value, b.value
    hp.init("YAML")
    want = {"envs": {"KUBERNETES_CDI_VENDOR_CLASS": "nvidia.com/gpu"}, "cdi_devices": ["nvidia.com/gpu=0", "nvidia.com/gpu=1", "nvidia.com/gpu=3"]}
    assert hp.allocate(["75", "214"]) == want and counts() == (0, 3)
    # the device is re-bound behind the plugin's back WITHOUT an event: the snapshot still answers (that is the
    # contract: the kernel announces every re-bind) ...
    link = os.path.join(base, "0000:c1:00.0", "iommu_group")
    os.unlink(link)
    os.symlink(os.path.join(str(tmp_path), "iommu_groups", "215"), link)
    assert hp.allocate(["214"])["cdi_devices"] == ["nvidia.com/gpu=3"] and counts() == (0, 4)
    # ... and with the event the live reads run and reject it exactly like the reference
    gen.value = 8
    with pytest.raises(RuntimeError, match="invalid allocation request: unknown device: 0000:c1:00.0"):
        hp.allocate(["214"])
    assert counts()[0] == 1
    assert hp.allocate(["75"])["cdi_devices"] == ["nvidia.com/gpu=0", "nvidia.com/gpu=1"] and counts()[0] == 3
    # an unhealthy watcher (lost messages) also means live reads
    gen.value, healthy.value = 7, 0
    assert hp.allocate(["76"])["cdi_devices"] == ["nvidia.com/gpu=2"] and counts()[0] == 4
    hp.close()
