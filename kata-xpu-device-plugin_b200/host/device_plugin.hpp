// device_plugin.hpp -- host side of the discovery hot path, above the C ABI (include/kxpu.h).
//
// The reference host is Go; this image has no Go toolchain, so the host logic that sits
// above libkxpu.so is written in C++ and mirrors the reference's package
// pkg/device_plugin one to one: same function names, same package-level seams
// (basePath, pciIdsFilePath, readLink, readIDFromFile -- device_plugin.go:36-39,
// returnIommuMap -- generic_device_plugin.go:34), same argument meaning, same error
// behaviour ("log and degrade": unreadable entries are skipped, an unknown device id
// falls back to the raw id, an Allocate re-validation failure is an error naming the bdf).
// The syscalls (walk, read, readlink) stay on the host exactly where the reference does
// them; everything per-device / per-byte after that goes through the kxpu_* calls.
// The gRPC server itself (Register / ListAndWatch stream / Allocate handler plumbing,
// generic_device_plugin.go:128-220) is out of scope and stays in the Go binary; the Go
// cgo shim that replaces this file in production is shown in INTEGRATION.md.
#pragma once
#include <cstdint>
#include <functional>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "../../include/kxpu.h"

namespace device_plugin {

// pkg/device_plugin/device_plugin.go:24-28
struct NvidiaGpuDevice {
    std::string addr;  // PCI address of device
    uint64_t index;    // PCI device index on PCI bus
};

// Go maps iterate in random order; the canonical order used here (and by the oracle) is
// first-seen walk order, which is one of the orders the reference can produce.
template <typename V>
using OrderedMap = std::vector<std::pair<std::string, V>>;

// pkg/device_plugin/generic_device_plugin.go:35-48 (the fields the hot path touches)
struct Device {  // pluginapi.Device
    std::string ID;
    std::string Health;
};
struct GenericDevicePlugin {
    std::string devpluginName;   // resource name suffix: "nvidia.com/<devpluginName>" (:211)
    std::string socketPath;      // DevicePluginPath + "kata-xpu-<name>.sock" (:76)
    std::string devicePath;      // "/dev/vfio/" (device_plugin.go:105)
    std::vector<Device> devs;
};

// pluginapi.ContainerAllocateResponse as Allocate fills it (generic_device_plugin.go:304-350)
struct ContainerAllocateResponse {
    std::map<std::string, std::string> Envs;
    std::vector<std::string> CDIDevices;  // CDIDevice{Name}
};

struct Error {
    bool failed = false;
    std::string message;
    explicit operator bool() const { return failed; }
};

// healthCheck (generic_device_plugin.go:389-457) + the re-send half of ListAndWatch (:222-250) for
// ONE plugin: an inotify watch on devicePath/<ID> of every device; Remove / Rename of the path marks
// the device Unhealthy, Create marks it Healthy.  The reference sends one ListAndWatchResponse per
// event; here poll() drains the whole queue, flips Health of the matching devs and reports how many
// changed, and the caller re-encodes the list ONCE per batch (Plugin::ListAndWatchBytes): the last
// response of a burst is byte-identical to the reference's last response.
// watchCreates = false mirrors the reference exactly: it only adds the device paths themselves (and the
// directory of the kubelet socket, which is out of scope here), so a Create of /dev/vfio/<group> is
// never seen.  watchCreates = true also watches devicePath itself, which is what the code intends.
class HealthWatcher {
  public:
    HealthWatcher(GenericDevicePlugin &dp, bool watchCreates = false);
    ~HealthWatcher();
    HealthWatcher(const HealthWatcher &) = delete;
    HealthWatcher &operator=(const HealthWatcher &) = delete;
    Error start();             // watcher.Add(devicePath/ID) for every dev (:421-430)
    int poll(int timeout_ms);  // #devs whose Health changed in this batch; 0 = none; -1 = error
    uint64_t events() const { return events_; }

  private:
    GenericDevicePlugin &dp_;
    bool watchCreates_;
    int fd_ = -1, dirWd_ = -1;
    std::map<int, std::string> wdToId_;
    uint64_t events_ = 0;
    int setHealth(const std::string &id, const char *health);
};

// SURVEY 8(f) row 2, second half: what makes a snapshot of the discovery trustworthy.  Allocate
// re-validates every device against sysfs (readlink iommu_group + read vendor, two syscalls per device,
// generic_device_plugin.go:329-338) because a device may have been re-bound since discovery.  The kernel
// announces exactly that: bind / unbind / add / remove uevents of the pci subsystem on the
// NETLINK_KOBJECT_UEVENT socket (bind / unbind since Linux 4.14).  BindWatcher counts them: while the
// generation it reports equals the one recorded at discovery, no PCI function changed its driver and the
// snapshot answers what the live reads would answer.
class BindWatcher {
  public:
    BindWatcher() = default;
    ~BindWatcher();
    BindWatcher(const BindWatcher &) = delete;
    BindWatcher &operator=(const BindWatcher &) = delete;
    Error start();            // opens the uevent socket (needs no privilege beyond a netlink socket)
    bool healthy() const { return fd_ >= 0 && !lost_; }
    // drains the socket; the generation grows by one per pci bind / unbind / add / remove event and jumps
    // when the kernel reports lost messages (ENOBUFS): then nothing can be said about what was missed
    uint64_t generation();
    // feeds one raw uevent message (tests; also what generation() calls per datagram)
    void feed(const char *msg, size_t len);

  private:
    int fd_ = -1;
    bool lost_ = false;
    uint64_t gen_ = 0;
};

class Plugin {
  public:
    // ---- seams (device_plugin.go:36-39, generic_device_plugin.go:34)
    std::string basePath = "/sys/bus/pci/devices";
    std::string pciIdsFilePath = "/usr/pci.ids";
    std::string cdiConfigPath = "/var/run/cdi/";  // device_plugin.go:20
    std::function<bool(const std::string &base, const std::string &addr, const std::string &link, std::string &out)> readLink;
    std::function<bool(const std::string &base, const std::string &addr, const std::string &prop, std::string &out)> readIDFromFile;
    std::function<const OrderedMap<std::vector<NvidiaGpuDevice>> &()> returnIommuMap;
    // Allocate re-validation (generic_device_plugin.go:329-338).  false (default) = the reference's live
    // reads for every device of every request.  true = answer from the discovery snapshot as long as
    // bindGeneration() still returns the value recorded by createIommuDeviceMap; any change (or an
    // unhealthy watcher) falls back to the live reads for that request.  bindGeneration is a seam: by
    // default it asks the BindWatcher (started on first use), tests replace it.
    bool snapshotValidation = false;
    std::function<bool(uint64_t &generation)> bindGeneration;
    uint64_t liveValidations = 0, snapshotValidations = 0;  // devices validated either way (tests, metrics)

    // ---- state (device_plugin.go:31,34)
    OrderedMap<std::vector<NvidiaGpuDevice>> iommuMap;  // group id -> devices
    OrderedMap<std::vector<std::string>> deviceMap;     // device id -> iommu groups
    std::vector<GenericDevicePlugin> devicePlugins;
    std::string lastCdiFile;

    explicit Plugin(kxpu_ctx *ctx);
    ~Plugin();

    // device_plugin.go:44-53 without the blocking gRPC part
    Error InitiateDevicePlugin();
    // device_plugin.go:126-180: walk + raw gather on the host, classify on the GPU (S1)
    Error createIommuDeviceMap();
    // device_plugin.go:208-259: parse-once table + batched lookup + sanitiser on the GPU (S2)
    std::string getDeviceName(const std::string &deviceID);
    // the same for a batch of ids: one kxpu_lookup + one kxpu_names (device_plugin.go:99 for every id of deviceMap)
    std::vector<std::string> getDeviceNames(const std::vector<std::string> &deviceIDs);
    // device_plugin.go:55-80 + cdi/spec.go:85-127: emit on the GPU, host writes the file (S3)
    Error generateCDISpec(const OrderedMap<std::vector<NvidiaGpuDevice>> &m, const std::string &format = "YAML");
    // device_plugin.go:83-112: per device id device lists + plugin objects (S4); nothing is started
    Error createDevicePlugins();
    // generic_device_plugin.go:320-355 for one container request (S5)
    Error Allocate(const std::vector<std::string> &devicesIDs, ContainerAllocateResponse &resp);
    // generic_device_plugin.go:224: the bytes of ListAndWatchResponse{Devices: dpi.devs}
    Error ListAndWatchBytes(const GenericDevicePlugin &dp, std::vector<uint8_t> &out);

    // raw gather only (no GPU): exposed for CPU tests of the walk
    Error gatherRecords(std::vector<kxpu_devrec> &recs);
    // the same records, read with openat / readlinkat relative to basePath by several threads
    // (SURVEY 8(f) row 2); falls back to gatherRecords when a seam was replaced.  threads = 0: automatic
    Error gatherRecordsFast(std::vector<kxpu_devrec> &recs, unsigned threads = 0);

  private:
    kxpu_ctx *ctx_;
    kxpu_table *table_ = nullptr;
    Error ensureTable();
    // first use: kxpu_pciids_join on pinned buffers = file -> table -> row handles of `keys` in one call
    Error loadAndJoin(const std::vector<uint32_t> &keys, std::vector<int32_t> &rows);
    BindWatcher bindWatcher_;
    bool haveSnapshotGen_ = false;
    uint64_t snapshotGen_ = 0;
};

}  // namespace device_plugin
