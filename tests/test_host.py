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
