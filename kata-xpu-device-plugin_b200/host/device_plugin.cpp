// device_plugin.cpp -- see device_plugin.hpp.  Reference: pkg/device_plugin/device_plugin.go,
// pkg/device_plugin/generic_device_plugin.go, cdi/spec.go (citations per function).
#include "device_plugin.hpp"

#include <dirent.h>
#include <poll.h>
#include <linux/netlink.h>
#include <sys/inotify.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <unistd.h>

#include <fcntl.h>

#include <algorithm>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

namespace device_plugin {

static const char *kDevicePluginPath = "/var/lib/kubelet/device-plugins/";  // pluginapi.DevicePluginPath
static const char *kHealthy = "Healthy";                                     // pluginapi.Healthy
static const char *kUnhealthy = "Unhealthy";                                 // pluginapi.Unhealthy
static const char *kK8SCDIVendorClass = "KUBERNETES_CDI_VENDOR_CLASS";       // generic_device_plugin.go:29
static const char *kCdiVendorClass = "nvidia.com/gpu";                       // generic_device_plugin.go:30

static Error fail(const std::string &m) { Error e; e.failed = true; e.message = m; return e; }
static Error kxfail(kxpu_ctx *ctx, const char *what, int32_t rc) {
    return fail(std::string(what) + ": " + kxpu_strerror(rc) + " (" + kxpu_last_error(ctx) + ")");
}

// readIDFromFileFunc, device_plugin.go:183-191 -- the os.ReadFile half.  Returns the RAW file
// bytes: the data[2:] / Trim("\n") half of :189 runs on the GPU (kxpu_classify).
static bool readIDFromFileFunc(const std::string &base, const std::string &addr, const std::string &prop, std::string &out) {
    std::string path = base + "/" + addr + "/" + prop;  // filepath.Join
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) {
        fprintf(stderr, "Could not read %s for device %s: %s\n", prop.c_str(), addr.c_str(), strerror(errno));
        return false;
    }
    char buf[256];
    size_t n = fread(buf, 1, sizeof buf, f);
    bool err = ferror(f) != 0;
    fclose(f);
    if (err) return false;
    out.assign(buf, n);
    return true;
}

// readLinkFunc, device_plugin.go:194-202: os.Readlink + last path element.
static bool readLinkFunc(const std::string &base, const std::string &addr, const std::string &link, std::string &out) {
    std::string path = base + "/" + addr + "/" + link;
    char buf[4096];
    ssize_t n = readlink(path.c_str(), buf, sizeof buf - 1);
    if (n < 0) {
        fprintf(stderr, "Could not read link %s for device %s: %s\n", link.c_str(), addr.c_str(), strerror(errno));
        return false;
    }
    std::string target(buf, (size_t)n);
    size_t slash = target.find_last_of('/');  // filepath.Split
    out = slash == std::string::npos ? target : target.substr(slash + 1);
    return true;
}

Plugin::Plugin(kxpu_ctx *ctx) : ctx_(ctx) {
    readLink = readLinkFunc;
    readIDFromFile = readIDFromFileFunc;
    returnIommuMap = [this]() -> const OrderedMap<std::vector<NvidiaGpuDevice>> & { return iommuMap; };
    bindGeneration = [this](uint64_t &generation) {
        if (!bindWatcher_.healthy() && bindWatcher_.start()) return false;
        generation = bindWatcher_.generation();
        return bindWatcher_.healthy();
    };
}

// ---------------------------------------------------------------------------- bind / unbind uevents
BindWatcher::~BindWatcher() {
    if (fd_ >= 0) close(fd_);
}

Error BindWatcher::start() {
    if (fd_ >= 0) close(fd_);
    lost_ = false;
    fd_ = socket(AF_NETLINK, SOCK_DGRAM | SOCK_CLOEXEC | SOCK_NONBLOCK, NETLINK_KOBJECT_UEVENT);
    if (fd_ < 0) return fail(std::string("uevent socket: ") + strerror(errno));
    struct sockaddr_nl sa;
    memset(&sa, 0, sizeof sa);
    sa.nl_family = AF_NETLINK;
    sa.nl_groups = 1;  // kernel uevent multicast group
    if (bind(fd_, (struct sockaddr *)&sa, sizeof sa) != 0) {
        Error e = fail(std::string("uevent bind: ") + strerror(errno));
        close(fd_);
        fd_ = -1;
        return e;
    }
    return Error();
}

// "ACTION@DEVPATH\0KEY=VALUE\0..." -- a pci function changing its driver or coming / going
void BindWatcher::feed(const char *msg, size_t len) {
    const char *end = msg + len;
    std::string action, subsystem;
    for (const char *p = msg; p < end; p += strlen(p) + 1) {
        if (strncmp(p, "ACTION=", 7) == 0) action = p + 7;
        else if (strncmp(p, "SUBSYSTEM=", 10) == 0) subsystem = p + 10;
        if (memchr(p, 0, (size_t)(end - p)) == nullptr) break;  // unterminated tail
    }
    if (subsystem == "pci" && (action == "bind" || action == "unbind" || action == "add" || action == "remove")) gen_++;
}

uint64_t BindWatcher::generation() {
    if (fd_ < 0) return gen_;
    char buf[8192];
    for (;;) {
        ssize_t k = recv(fd_, buf, sizeof buf - 1, 0);
        if (k > 0) {
            buf[k] = 0;
            feed(buf, (size_t)k);
            continue;
        }
        if (k < 0 && errno == EINTR) continue;
        if (k < 0 && errno == ENOBUFS) { lost_ = true; gen_ += 1ull << 32; }  // messages were dropped
        break;
    }
    return gen_;
}

Plugin::~Plugin() {
    if (table_) kxpu_table_free(ctx_, table_);
}

// the Trim half of readIDFromFileFunc (:189): data[2:] with '\n' trimmed at both ends
static std::string trimID(const std::string &raw) {
    if (raw.size() < 2) return std::string();
    size_t a = 2, b = raw.size();
    while (a < b && raw[a] == '\n') a++;
    while (b > a && raw[b - 1] == '\n') b--;
    return raw.substr(a, b - a);
}

// an id file as the record carries it: the raw bytes when they fit the 8-byte field, else the
// canonical spelling "0x" + id + "\n" of the same id (identical after the reference's data[2:] /
// Trim); false when even that does not fit (the id itself is longer than 5 characters)
static bool packID(const std::string &raw, uint8_t txt[8], uint8_t &len) {
    if (raw.size() <= 8) {
        memcpy(txt, raw.data(), raw.size());
        len = (uint8_t)raw.size();
        return true;
    }
    const std::string id = trimID(raw);
    if (id.size() > 5 || id.find('\n') != std::string::npos) return false;
    const std::string canon = "0x" + id + "\n";
    memcpy(txt, canon.data(), canon.size());
    len = (uint8_t)canon.size();
    return true;
}

// The body of the walk callback for one non-directory entry (device_plugin.go:141-175, the reads
// only, in the reference's order and as lazily as the reference: nothing is read behind a vendor
// that is not 10de or a driver that is not vfio-pci): raw bytes of `vendor` / `device`, basenames
// of the `driver` / `iommu_group` links.  An entry the record cannot carry (address longer than 15
// bytes, group that is not a canonical decimal below 2^32-1, id longer than the field) is logged and
// skipped like a read error -- and only if the reference would have accepted it; it never stops the walk.
template <typename ReadID, typename ReadLnk>
static Error leafRecord(const std::string &name, ReadID readID, ReadLnk readLnk, kxpu_devrec &r) {
    memset(&r, 0, sizeof r);
    strncpy(r.bdf, name.c_str(), sizeof r.bdf - 1);
    std::string s;
    if (!readID("vendor", s)) {
        r.flags |= KXPU_REC_VENDOR_ERR;  // "Could not get vendor ID for device" -> skipped (:143-146)
        return Error();
    }
    const bool nvidia = trimID(s) == "10de";  // :149
    if (!packID(s, r.vendor_txt, r.vendor_len)) {  // an id of six or more characters is not 10de
        memcpy(r.vendor_txt, s.data(), 8);
        r.vendor_len = 8;
        r.flags |= KXPU_REC_VENDOR_ERR;
    }
    if (!nvidia) return Error();
    if (!readLnk("driver", s)) {
        r.flags |= KXPU_REC_DRIVER_ERR;  // :152-155
        return Error();
    }
    memcpy(r.driver, s.data(), std::min<size_t>(s.size(), sizeof r.driver - 1));
    if (s != "vfio-pci") return Error();  // :156
    if (name.size() > sizeof r.bdf - 1) {
        fprintf(stderr, "PCI address longer than 15 bytes, device skipped: %s\n", name.c_str());
        r.flags |= KXPU_REC_IOMMU_ERR;
        return Error();
    }
    if (readLnk("iommu_group", s)) {  // :157
        bool dec = !s.empty() && s.size() <= 10;
        unsigned long long v = 0;
        for (char c : s) { if (c < '0' || c > '9') dec = false; else v = v * 10 + (unsigned)(c - '0'); }
        if (!dec || v >= 0xFFFFFFFFull || (s.size() > 1 && s[0] == '0')) {
            fprintf(stderr, "iommu_group of %s is not a canonical decimal number below 2^32-1, device skipped: %s\n", name.c_str(), s.c_str());
            r.flags |= KXPU_REC_IOMMU_ERR;
            return Error();
        }
        r.iommu_group = (uint32_t)v;
    } else {
        r.flags |= KXPU_REC_IOMMU_ERR;  // :158-161
        return Error();
    }
    // :164 reads `device` only for the first member of a group; which record that is is decided on
    // the GPU, so the file is read for every accepted candidate (a failure only matters for a first member)
    if (readID("device", s)) {
        if (!packID(s, r.device_txt, r.device_len)) {
            fprintf(stderr, "device id of %s is longer than the record field, device skipped\n", name.c_str());
            r.flags |= KXPU_REC_DEVICE_ERR;
        }
    } else {
        r.flags |= KXPU_REC_DEVICE_ERR;  // :165-168
    }
    return Error();
}

// filepath.Walk(basePath, ...) (device_plugin.go:132): lexical order, os.Lstat (symlinks are not
// followed, so a real sysfs entry is "not a directory"), directories are descended into and
// reported as "Not a device" (:137-140).
static Error walkDir(Plugin &p, const std::string &path, const std::string &name, std::vector<kxpu_devrec> &recs) {
    struct stat sb;
    if (lstat(path.c_str(), &sb) != 0) return fail("Error accessing file path \"" + path + "\": " + strerror(errno));  // :133-136
    if (S_ISDIR(sb.st_mode)) {
        DIR *d = opendir(path.c_str());
        if (!d) return fail("Error accessing file path \"" + path + "\": " + strerror(errno));
        std::vector<std::string> names;
        while (struct dirent *de = readdir(d)) {
            if (strcmp(de->d_name, ".") == 0 || strcmp(de->d_name, "..") == 0) continue;
            names.push_back(de->d_name);
        }
        closedir(d);
        std::sort(names.begin(), names.end());
        for (const std::string &n : names) {
            Error e = walkDir(p, path + "/" + n, n, recs);
            if (e) return e;
        }
        return Error();
    }
    // one raw record per non-directory entry; every read goes through the seams and is keyed by
    // info.Name() under basePath exactly like :142,:151,:157,:164
    kxpu_devrec r;
    Error e = leafRecord(name,
                         [&](const char *prop, std::string &out) { return p.readIDFromFile(p.basePath, name, prop, out); },
                         [&](const char *link, std::string &out) { return p.readLink(p.basePath, name, link, out); }, r);
    if (e) return e;
    recs.push_back(r);
    return Error();
}

Error Plugin::gatherRecords(std::vector<kxpu_devrec> &recs) {
    recs.clear();
    size_t slash = basePath.find_last_of('/');
    return walkDir(*this, basePath, slash == std::string::npos ? basePath : basePath.substr(slash + 1), recs);
}

// ---------------------------------------------------------------------------- SURVEY 8(f) row 2
// Batched sysfs ingestion: the same records as gatherRecords, but the entries of basePath are read
// with paths RELATIVE to one directory descriptor (openat / readlinkat on "<bdf>/vendor": no lstat per
// entry -- getdents64 already says what is a directory --, no absolute path resolution, no FILE
// buffering) and by several threads, each filling its own slice of the record table, so that S1 ends
// in one contiguous table ready for a single H2D copy.  Only with the default seams; real directories
// under basePath (never on sysfs) go through the generic walk at their position.
Error Plugin::gatherRecordsFast(std::vector<kxpu_devrec> &recs, unsigned threads) {
    recs.clear();
    struct stat sb;
    if (lstat(basePath.c_str(), &sb) != 0) return fail("Error accessing file path \"" + basePath + "\": " + strerror(errno));
    // the fast reads bypass the seams: only when nobody replaced them (tests do, device_plugin.go:38-39)
    using SeamFn = bool (*)(const std::string &, const std::string &, const std::string &, std::string &);
    SeamFn const *rl = readLink.target<SeamFn>(), *ri = readIDFromFile.target<SeamFn>();
    const bool defaultSeams = rl && *rl == readLinkFunc && ri && *ri == readIDFromFileFunc;
    if (!S_ISDIR(sb.st_mode) || !defaultSeams) return gatherRecords(recs);
    int basefd = open(basePath.c_str(), O_RDONLY | O_DIRECTORY | O_CLOEXEC);
    if (basefd < 0) return fail("Error accessing file path \"" + basePath + "\": " + strerror(errno));
    DIR *d = fdopendir(dup(basefd));
    if (!d) { close(basefd); return fail("Error accessing file path \"" + basePath + "\": " + strerror(errno)); }
    struct Ent { std::string name; bool dir; };
    std::vector<Ent> ents;
    while (struct dirent *de = readdir(d)) {
        if (strcmp(de->d_name, ".") == 0 || strcmp(de->d_name, "..") == 0) continue;
        bool isdir = de->d_type == DT_DIR;
        if (de->d_type == DT_UNKNOWN) {  // file systems without d_type: one fstatat, still no path walk
            struct stat es;
            isdir = fstatat(basefd, de->d_name, &es, AT_SYMLINK_NOFOLLOW) == 0 && S_ISDIR(es.st_mode);
        }
        ents.push_back(Ent{de->d_name, isdir});
    }
    closedir(d);
    std::sort(ents.begin(), ents.end(), [](const Ent &a, const Ent &b) { return a.name < b.name; });

    const size_t N = ents.size();
    std::vector<kxpu_devrec> flat(N);
    std::vector<Error> errs(N);
    auto readID = [&](const std::string &name, const char *prop, std::string &out) {
        const std::string rel = name + "/" + prop;
        int fd = openat(basefd, rel.c_str(), O_RDONLY | O_CLOEXEC);
        if (fd < 0) {
            fprintf(stderr, "Could not read %s for device %s: %s\n", prop, name.c_str(), strerror(errno));
            return false;
        }
        char buf[256];
        size_t got = 0;
        for (;;) {  // os.ReadFile reads to EOF
            ssize_t k = read(fd, buf + got, sizeof buf - got);
            if (k < 0) { close(fd); return false; }
            if (k == 0) break;
            got += (size_t)k;
            if (got == sizeof buf) break;
        }
        close(fd);
        out.assign(buf, got);
        return true;
    };
    auto readLnk = [&](const std::string &name, const char *link, std::string &out) {
        const std::string rel = name + "/" + link;
        char buf[4096];
        ssize_t k = readlinkat(basefd, rel.c_str(), buf, sizeof buf - 1);
        if (k < 0) {
            fprintf(stderr, "Could not read link %s for device %s: %s\n", link, name.c_str(), strerror(errno));
            return false;
        }
        std::string target(buf, (size_t)k);
        size_t slash = target.find_last_of('/');
        out = slash == std::string::npos ? target : target.substr(slash + 1);
        return true;
    };
    auto work = [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; i++) {
            if (ents[i].dir) continue;
            const std::string &name = ents[i].name;
            errs[i] = leafRecord(name, [&](const char *prop, std::string &out) { return readID(name, prop, out); },
                                 [&](const char *link, std::string &out) { return readLnk(name, link, out); }, flat[i]);
        }
    };
    if (threads == 0) threads = std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
    threads = (unsigned)std::min<size_t>(threads, std::max<size_t>(N / 64, 1));
    if (threads <= 1) {
        work(0, N);
    } else {
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < threads; t++) pool.emplace_back(work, N * t / threads, N * (t + 1) / threads);
        for (auto &th : pool) th.join();
    }
    close(basefd);
    // assemble in walk order; the first error in walk order wins, like the sequential walk
    for (size_t i = 0; i < N; i++) {
        if (ents[i].dir) {
            Error e = walkDir(*this, basePath + "/" + ents[i].name, ents[i].name, recs);
            if (e) return e;
        } else {
            if (errs[i]) return errs[i];
            recs.push_back(flat[i]);
        }
    }
    return Error();
}

static std::string devIdString(uint64_t packed) {
    char b[9];
    memcpy(b, &packed, 8);
    b[8] = 0;
    return std::string(b);
}

// createIommuDeviceMap, device_plugin.go:126-180
Error Plugin::createIommuDeviceMap() {
    iommuMap.clear();   // :127
    deviceMap.clear();  // :128
    // the generation is read BEFORE the walk: an event during the walk makes the snapshot stale, never fresh
    haveSnapshotGen_ = snapshotValidation && bindGeneration && bindGeneration(snapshotGen_);
    std::vector<kxpu_devrec> recs;
    Error e = gatherRecordsFast(recs);  // same records as gatherRecords (falls back to it when a seam was replaced)
    if (e) { fprintf(stderr, "%s\n", e.message.c_str()); }  // Walk's error is ignored by the reference (:132)
    const size_t n = recs.size();
    std::vector<uint32_t> accept(n), gids(n), goff(n + 1), gmem(n), doff(n + 1), dgrp(n);
    std::vector<uint64_t> dids(n);
    kxpu_classify_out out;
    memset(&out, 0, sizeof out);
    out.accept_index = accept.data(); out.group_ids = gids.data(); out.group_off = goff.data();
    out.group_members = gmem.data(); out.dev_ids = dids.data(); out.dev_off = doff.data(); out.dev_groups = dgrp.data();
    int32_t rc = kxpu_classify(ctx_, recs.data(), n, &out);
    if (rc != KXPU_OK) return kxfail(ctx_, "kxpu_classify", rc);  // fatal: there is no CPU path
    for (uint32_t g = 0; g < out.n_groups; g++) {
        std::vector<NvidiaGpuDevice> devs;
        for (uint32_t k = goff[g]; k < goff[g + 1]; k++) {
            uint32_t i = gmem[k];
            devs.push_back(NvidiaGpuDevice{std::string(recs[i].bdf), accept[i]});  // :171-174
        }
        iommuMap.emplace_back(std::to_string(gids[g]), std::move(devs));
    }
    for (uint32_t d = 0; d < out.n_devids; d++) {
        std::vector<std::string> groups;
        for (uint32_t k = doff[d]; k < doff[d + 1]; k++) groups.push_back(std::to_string(dgrp[k]));  // :169
        deviceMap.emplace_back(devIdString(dids[d]), std::move(groups));
    }
    return Error();
}

// The pci.ids file into page-locked memory the GPU can address (kxpu_pinned_alloc): with text, keys and rows in
// such buffers kxpu_pciids_join copies nothing -- one cooperative kernel pulls the text over PCIe, builds the
// table and writes the row handles back (include/kxpu.h).  Falls back to ordinary memory when pinning fails.
namespace {
struct PinnedBuf {
    kxpu_ctx *ctx;
    void *p = nullptr;
    bool pinned = false;
    size_t n = 0;
    PinnedBuf(kxpu_ctx *c, size_t bytes) : ctx(c), n(bytes) {
        if (kxpu_pinned_alloc(ctx, bytes ? bytes : 16, &p) == KXPU_OK && p) pinned = true;
        else p = malloc(bytes ? bytes : 16);
    }
    ~PinnedBuf() {
        if (pinned) kxpu_pinned_free(ctx, p);
        else free(p);
    }
    PinnedBuf(const PinnedBuf &) = delete;
    PinnedBuf &operator=(const PinnedBuf &) = delete;
};
}  // namespace

// First use: ONE kxpu_pciids_join call reads the file (device_plugin.go:210), builds the table and joins `keys`
// (the start-up batch of createDevicePlugins, device_plugin.go:91-105).  keys may be empty (load only).
Error Plugin::loadAndJoin(const std::vector<uint32_t> &keys, std::vector<int32_t> &rows) {
    FILE *f = fopen(pciIdsFilePath.c_str(), "rb");  // device_plugin.go:210
    if (!f) return fail("Error opening pci ids file " + pciIdsFilePath);
    std::vector<uint8_t> head;
    size_t n = 0;
    if (fseek(f, 0, SEEK_END) == 0) {
        const long sz = ftell(f);
        if (sz > 0) n = (size_t)sz;
        rewind(f);
    }
    PinnedBuf text(ctx_, n), hk(ctx_, keys.size() * 4), hr(ctx_, keys.size() * 4);
    if (!text.p || !hk.p || !hr.p) { fclose(f); return fail("out of memory reading " + pciIdsFilePath); }
    const size_t got = n ? fread(text.p, 1, n, f) : 0;
    fclose(f);
    if (!keys.empty()) memcpy(hk.p, keys.data(), keys.size() * 4);
    const int32_t rc = kxpu_pciids_join(ctx_, (const uint8_t *)text.p, got, (const uint32_t *)hk.p, keys.size(), (int32_t *)hr.p, &table_);
    if (rc != KXPU_OK) { table_ = nullptr; return kxfail(ctx_, "kxpu_pciids_join", rc); }
    rows.assign((const int32_t *)hr.p, (const int32_t *)hr.p + keys.size());
    return Error();
}

Error Plugin::ensureTable() {
    if (table_) return Error();
    std::vector<int32_t> none;
    return loadAndJoin({}, none);
}

static bool parseHex4(const std::string &s, uint32_t &v) {
    if (s.size() != 4) return false;
    v = 0;
    for (char c : s) {
        uint32_t d;
        if (c >= '0' && c <= '9') d = (uint32_t)(c - '0');
        else if (c >= 'a' && c <= 'f') d = (uint32_t)(c - 'a' + 10);
        else return false;
        v = v * 16 + d;
    }
    return true;
}

// getDeviceName, device_plugin.go:208-259, for a whole batch of device ids: ONE join and ONE name
// gather through the ABI (cgo calls stay coarse).  "" means "not found" exactly like the
// reference; sysfs ids are four lowercase hex digits, anything else is treated as not found.
std::vector<std::string> Plugin::getDeviceNames(const std::vector<std::string> &deviceIDs) {
    std::vector<std::string> out(deviceIDs.size());
    std::vector<uint32_t> keys;
    std::vector<size_t> where;
    for (size_t i = 0; i < deviceIDs.size(); i++) {
        uint32_t d;
        if (parseHex4(deviceIDs[i], d)) { keys.push_back((0x10deu << 16) | d); where.push_back(i); }  // nvidiaVendorID, :19
    }
    std::vector<int32_t> rows(keys.size(), KXPU_ROW_MISS);
    if (!table_) {
        // start-up: file -> table -> row handles of the whole batch in one call
        Error e = loadAndJoin(keys, rows);
        if (e) { fprintf(stderr, "%s\n", e.message.c_str()); return out; }  // :211-214
        if (keys.empty()) return out;
    } else {
        if (keys.empty()) return out;
        if (kxpu_lookup(ctx_, table_, keys.data(), keys.size(), rows.data()) != KXPU_OK) return out;
    }
    std::vector<uint32_t> offs(keys.size() + 1);
    size_t need = 0;
    int32_t rc = kxpu_names(ctx_, table_, rows.data(), rows.size(), nullptr, 0, offs.data(), &need);
    if (rc != KXPU_OK && rc != KXPU_E_NOSPACE) return out;
    std::vector<uint8_t> blob(need ? need : 1);
    if (need && kxpu_names(ctx_, table_, rows.data(), rows.size(), blob.data(), need, offs.data(), &need) != KXPU_OK) return out;
    for (size_t k = 0; k < keys.size(); k++) {
        if (rows[k] == KXPU_ROW_MISS) {
            fprintf(stderr, "Could not find NVIDIA device with id: %s\n", deviceIDs[where[k]].c_str());  // :234
            continue;
        }
        out[where[k]].assign((const char *)blob.data() + offs[k], offs[k + 1] - offs[k]);
    }
    return out;
}

std::string Plugin::getDeviceName(const std::string &deviceID) { return getDeviceNames({deviceID})[0]; }

// generateCDISpec, device_plugin.go:55-80 + CdiSpec.Save, cdi/spec.go:85-127
Error Plugin::generateCDISpec(const OrderedMap<std::vector<NvidiaGpuDevice>> &m, const std::string &format) {
    std::vector<kxpu_cdidev> devs;
    for (const auto &kv : m) {
        for (const NvidiaGpuDevice &dev : kv.second) {
            kxpu_cdidev c;
            memset(&c, 0, sizeof c);
            strncpy(c.bdf, dev.addr.c_str(), sizeof c.bdf - 1);
            c.iommu_group = (uint32_t)strtoul(kv.first.c_str(), nullptr, 10);
            c.index = dev.index;
            devs.push_back(c);
        }
    }
    // Go ranges over the map in random order (:59); canonical order = ascending index
    std::sort(devs.begin(), devs.end(), [](const kxpu_cdidev &a, const kxpu_cdidev &b) { return a.index < b.index; });
    const int32_t fmt = format == "YAML" ? KXPU_FMT_YAML : KXPU_FMT_JSON;  // spec.go:86-89,102-114
    size_t len = 0;
    int32_t rc = kxpu_cdi_emit(ctx_, fmt, devs.data(), devs.size(), nullptr, 0, &len);
    if (rc != KXPU_OK && rc != KXPU_E_NOSPACE) return kxfail(ctx_, "kxpu_cdi_emit", rc);
    std::vector<uint8_t> doc(len ? len : 1);
    rc = kxpu_cdi_emit(ctx_, fmt, devs.data(), devs.size(), doc.data(), len, &len);
    if (rc != KXPU_OK) return kxfail(ctx_, "kxpu_cdi_emit", rc);
    const std::string file_path = cdiConfigPath + "cdi-vfio-xxxx" + (fmt == KXPU_FMT_YAML ? ".yaml" : ".json");  // :79, spec.go:92
    FILE *f = fopen(file_path.c_str(), "wb");  // os.Create
    if (!f) {
        printf("Error creating file: %s\n", strerror(errno));  // spec.go:95: printed and swallowed
        return Error();
    }
    size_t w = fwrite(doc.data(), 1, len, f);
    fclose(f);
    if (w != len) { printf("Error writing to file\n"); return Error(); }
    lastCdiFile = file_path;
    printf("Data successfully written to file\n");  // spec.go:126
    return Error();
}

// createDevicePlugins, device_plugin.go:83-112 (nothing is started: no gRPC here)
Error Plugin::createDevicePlugins() {
    devicePlugins.clear();
    std::vector<std::string> ids;
    for (const auto &kv : deviceMap) ids.push_back(kv.first);
    const std::vector<std::string> names = getDeviceNames(ids);  // :99 for every device id at once
    size_t at = 0;
    for (const auto &kv : deviceMap) {  // :91
        GenericDevicePlugin dp;
        for (const std::string &dev : kv.second) dp.devs.push_back(Device{dev, kHealthy});  // :93-98
        std::string devpluginName = names[at++];
        if (devpluginName.empty()) {
            fprintf(stderr, "Error: Could not find device name for device id: %s\n", kv.first.c_str());
            devpluginName = kv.first;  // :100-103
        }
        dp.devpluginName = devpluginName;
        dp.devicePath = "/dev/vfio/";                                                          // :105
        dp.socketPath = std::string(kDevicePluginPath) + "kata-xpu-" + devpluginName + ".sock";  // generic:76
        devicePlugins.push_back(std::move(dp));
    }
    return Error();
}

Error Plugin::InitiateDevicePlugin() {
    Error e = createIommuDeviceMap();  // :46
    if (e) return e;
    e = generateCDISpec(iommuMap);  // :49
    if (e) return e;
    return createDevicePlugins();  // :52
}

// Allocate, generic_device_plugin.go:320-355, for one ContainerAllocateRequest
Error Plugin::Allocate(const std::vector<std::string> &devicesIDs, ContainerAllocateResponse &resp) {
    std::vector<uint64_t> devIndexes;
    const auto &returnedMap = returnIommuMap();
    // snapshot validation (off by default): every device of returnedMap was NVIDIA, bound to vfio-pci and in
    // this group when it was discovered; if no pci function was bound / unbound / added / removed since, the
    // live reads below would return exactly that
    bool fromSnapshot = false;
    if (snapshotValidation && haveSnapshotGen_ && bindGeneration) {
        uint64_t now = 0;
        fromSnapshot = bindGeneration(now) && now == snapshotGen_;
    }
    for (const std::string &iommuId : devicesIDs) {  // :324
        const std::vector<NvidiaGpuDevice> *nvDevs = nullptr;
        for (const auto &kv : returnedMap) if (kv.first == iommuId) { nvDevs = &kv.second; break; }
        if (!nvDevs) continue;  // unknown group id: empty nvDevs, no error (:327)
        for (const NvidiaGpuDevice &dev : *nvDevs) {
            if (fromSnapshot) {
                snapshotValidations++;
                devIndexes.push_back(dev.index);  // :340
                continue;
            }
            liveValidations++;
            std::string iommuGroup, vendor;
            if (!readLink(basePath, dev.addr, "iommu_group", iommuGroup) || iommuGroup != iommuId)  // :329-333
                return fail("invalid allocation request: unknown device: " + dev.addr);
            if (!readIDFromFile(basePath, dev.addr, "vendor", vendor) || trimID(vendor) != "10de")  // :334-338
                return fail("invalid allocation request: unknown device: " + dev.addr);
            devIndexes.push_back(dev.index);  // :340
        }
    }
    resp.CDIDevices.clear();
    if (!devIndexes.empty()) {  // updateResponseForCDI :274-299; strategy cdi-cri is on (:61)
        std::vector<uint32_t> offs(devIndexes.size() + 1);
        std::vector<uint8_t> buf(36 * devIndexes.size());
        size_t need = 0;
        int32_t rc = kxpu_alloc_names(ctx_, devIndexes.data(), devIndexes.size(), buf.data(), buf.size(), offs.data(), &need);
        if (rc != KXPU_OK) return fail("failed to get allocate response: " + std::string(kxpu_strerror(rc)));
        for (size_t i = 0; i < devIndexes.size(); i++)
            resp.CDIDevices.emplace_back((const char *)buf.data() + offs[i], offs[i + 1] - offs[i]);
    }
    resp.Envs.clear();
    resp.Envs[kK8SCDIVendorClass] = kCdiVendorClass;  // :348-350 overwrites Envs
    return Error();
}

Error Plugin::ListAndWatchBytes(const GenericDevicePlugin &dp, std::vector<uint8_t> &out) {
    std::vector<uint32_t> groups;
    std::vector<uint8_t> healthy;
    for (const Device &d : dp.devs) {
        groups.push_back((uint32_t)strtoul(d.ID.c_str(), nullptr, 10));
        healthy.push_back(d.Health == kHealthy);
    }
    size_t len = 0;
    int32_t rc = kxpu_lw_encode(ctx_, groups.data(), healthy.data(), groups.size(), nullptr, 0, &len);
    if (rc != KXPU_OK && rc != KXPU_E_NOSPACE) return kxfail(ctx_, "kxpu_lw_encode", rc);
    out.resize(len);
    if (len == 0) return Error();
    rc = kxpu_lw_encode(ctx_, groups.data(), healthy.data(), groups.size(), out.data(), len, &len);
    if (rc != KXPU_OK) return kxfail(ctx_, "kxpu_lw_encode", rc);
    return Error();
}

// ---------------------------------------------------------------------------- health (row 3 of SURVEY 8(f))
HealthWatcher::HealthWatcher(GenericDevicePlugin &dp, bool watchCreates) : dp_(dp), watchCreates_(watchCreates) {}

HealthWatcher::~HealthWatcher() {
    if (fd_ >= 0) close(fd_);
}

// filepath.Join(path, dev.ID): one separator, no trailing one
static std::string joinPath(const std::string &dir, const std::string &name) {
    if (dir.empty()) return name;
    return dir.back() == '/' ? dir + name : dir + "/" + name;
}

Error HealthWatcher::start() {
    fd_ = inotify_init1(IN_NONBLOCK | IN_CLOEXEC);  // fsnotify.NewWatcher (:396)
    if (fd_ < 0) return fail(std::string("Unable to create fsnotify watcher: ") + strerror(errno));
    // fsnotify's inotify backend adds every path with this mask; only the ops healthCheck looks at matter
    const uint32_t mask = IN_DELETE_SELF | IN_MOVE_SELF | IN_ATTRIB | IN_MODIFY;
    for (const Device &dev : dp_.devs) {  // :421-430
        const std::string devicePath = joinPath(dp_.devicePath, dev.ID);
        int wd = inotify_add_watch(fd_, devicePath.c_str(), mask);
        if (wd < 0) return fail("Unable to add device path to fsnotify watcher: " + devicePath + ": " + strerror(errno));
        wdToId_[wd] = dev.ID;
    }
    if (watchCreates_) {
        dirWd_ = inotify_add_watch(fd_, dp_.devicePath.c_str(), IN_CREATE | IN_MOVED_TO);
        if (dirWd_ < 0) return fail("Unable to add device directory to fsnotify watcher: " + dp_.devicePath + ": " + strerror(errno));
    }
    return Error();
}

// ListAndWatch's loops over dpi.devs (:230-234, :239-243): every dev with that ID
int HealthWatcher::setHealth(const std::string &id, const char *health) {
    int changed = 0;
    for (Device &d : dp_.devs)
        if (d.ID == id && d.Health != health) { d.Health = health; changed++; }
    return changed;
}

int HealthWatcher::poll(int timeout_ms) {
    if (fd_ < 0) return -1;
    struct pollfd pfd = {fd_, POLLIN, 0};
    int pr = ::poll(&pfd, 1, timeout_ms);
    if (pr < 0) return errno == EINTR ? 0 : -1;
    if (pr == 0) return 0;
    int changed = 0;
    alignas(struct inotify_event) char buf[16384];
    for (;;) {
        ssize_t n = read(fd_, buf, sizeof buf);
        if (n <= 0) break;  // EAGAIN: queue drained
        for (char *p = buf; p < buf + n;) {
            const struct inotify_event *ev = reinterpret_cast<const struct inotify_event *>(p);
            p += sizeof(struct inotify_event) + ev->len;
            events_++;
            if (ev->wd == dirWd_ && ev->len > 0) {
                // fsnotify.Create of devicePath/<name> (:441-443)
                const std::string name(ev->name);
                for (const Device &d : dp_.devs)
                    if (d.ID == name) {
                        changed += setHealth(name, kHealthy);
                        // the old watch died with the old inode: watch the new file like a restarted healthCheck would
                        int wd = inotify_add_watch(fd_, joinPath(dp_.devicePath, name).c_str(),
                                                   IN_DELETE_SELF | IN_MOVE_SELF | IN_ATTRIB | IN_MODIFY);
                        if (wd >= 0) wdToId_[wd] = name;
                        break;
                    }
                continue;
            }
            auto it = wdToId_.find(ev->wd);
            if (it == wdToId_.end()) continue;
            if (ev->mask & (IN_DELETE_SELF | IN_MOVE_SELF))  // fsnotify.Remove / fsnotify.Rename (:444-448)
                changed += setHealth(it->second, kUnhealthy);
            if (ev->mask & IN_IGNORED) wdToId_.erase(it);   // the kernel dropped the watch (file gone)
        }
    }
    return changed;
}

}  // namespace device_plugin

// ----------------------------------------------------------------------------------------
// C surface for the Python tests (tests/test_host*.py): drives the class above the way
// cmd/main.go drives the Go package.
// ----------------------------------------------------------------------------------------
using device_plugin::Plugin;

static void jstr(std::string &o, const std::string &s) {
    o += '"';
    for (char c : s) { if (c == '"' || c == '\\') o += '\\'; o += c; }
    o += '"';
}

static int copy_out(const std::string &s, char *out, size_t cap) {
    if (s.size() + 1 > cap) return -1;
    memcpy(out, s.c_str(), s.size() + 1);
    return (int)s.size();
}

extern "C" {

// CPU only: the raw gather of createIommuDeviceMap (no GPU involved)
int kxh_gather(const char *base_path, kxpu_devrec *out, size_t cap, size_t *n, char *err, size_t errcap) {
    Plugin p(nullptr);
    p.basePath = base_path;
    std::vector<kxpu_devrec> recs;
    device_plugin::Error e = p.gatherRecords(recs);
    if (e) { copy_out(e.message, err, errcap); return -1; }
    *n = recs.size();
    if (recs.size() > cap) return -2;
    memcpy(out, recs.data(), recs.size() * sizeof(kxpu_devrec));
    return 0;
}

// the batched / threaded variant of the same gather
int kxh_gather_fast(const char *base_path, unsigned threads, kxpu_devrec *out, size_t cap, size_t *n, char *err, size_t errcap) {
    Plugin p(nullptr);
    p.basePath = base_path;
    std::vector<kxpu_devrec> recs;
    device_plugin::Error e = p.gatherRecordsFast(recs, threads);
    if (e) { copy_out(e.message, err, errcap); return -1; }
    *n = recs.size();
    if (recs.size() > cap) return -2;
    memcpy(out, recs.data(), recs.size() * sizeof(kxpu_devrec));
    return 0;
}

void *kxh_new(kxpu_ctx *ctx, const char *base_path, const char *pciids_path, const char *cdi_dir) {
    Plugin *p = new Plugin(ctx);
    p->basePath = base_path;
    p->pciIdsFilePath = pciids_path;
    p->cdiConfigPath = cdi_dir;
    return p;
}
void kxh_free(void *h) { delete (Plugin *)h; }

// InitiateDevicePlugin + a JSON dump of the resulting state
int kxh_init(void *h, const char *format, char *json, size_t cap) {
    Plugin *p = (Plugin *)h;
    device_plugin::Error e = p->createIommuDeviceMap();
    if (!e) e = p->generateCDISpec(p->iommuMap, format);
    if (!e) e = p->createDevicePlugins();
    if (e) { copy_out(e.message, json, cap); return -1; }
    std::string o = "{\"iommuMap\":[";
    bool first = true;
    for (const auto &kv : p->iommuMap) {
        if (!first) o += ',';
        first = false;
        o += '['; jstr(o, kv.first); o += ",[";
        for (size_t i = 0; i < kv.second.size(); i++) {
            if (i) o += ',';
            o += '['; jstr(o, kv.second[i].addr); o += ',' + std::to_string(kv.second[i].index) + ']';
        }
        o += "]]";
    }
    o += "],\"deviceMap\":[";
    first = true;
    for (const auto &kv : p->deviceMap) {
        if (!first) o += ',';
        first = false;
        o += '['; jstr(o, kv.first); o += ",[";
        for (size_t i = 0; i < kv.second.size(); i++) { if (i) o += ','; jstr(o, kv.second[i]); }
        o += "]]";
    }
    o += "],\"plugins\":[";
    first = true;
    for (const auto &dp : p->devicePlugins) {
        if (!first) o += ',';
        first = false;
        o += "{\"name\":"; jstr(o, dp.devpluginName);
        o += ",\"resource\":"; jstr(o, "nvidia.com/" + dp.devpluginName);
        o += ",\"socket\":"; jstr(o, dp.socketPath);
        o += ",\"devs\":[";
        for (size_t i = 0; i < dp.devs.size(); i++) {
            if (i) o += ',';
            o += '['; jstr(o, dp.devs[i].ID); o += ','; jstr(o, dp.devs[i].Health); o += ']';
        }
        o += "]}";
    }
    o += "],\"cdiFile\":"; jstr(o, p->lastCdiFile); o += '}';
    return copy_out(o, json, cap);
}

// Allocate for one container request; ids = comma separated IOMMU group ids
int kxh_allocate(void *h, const char *ids_csv, char *json, size_t cap) {
    Plugin *p = (Plugin *)h;
    std::vector<std::string> ids;
    std::string cur;
    for (const char *c = ids_csv; *c; c++) { if (*c == ',') { ids.push_back(cur); cur.clear(); } else cur += *c; }
    if (!cur.empty() || (ids_csv[0] && ids_csv[strlen(ids_csv) - 1] == ',')) ids.push_back(cur);
    device_plugin::ContainerAllocateResponse resp;
    device_plugin::Error e = p->Allocate(ids, resp);
    if (e) { copy_out(e.message, json, cap); return -1; }
    std::string o = "{\"envs\":{";
    bool first = true;
    for (const auto &kv : resp.Envs) { if (!first) o += ','; first = false; jstr(o, kv.first); o += ':'; jstr(o, kv.second); }
    o += "},\"cdi_devices\":[";
    for (size_t i = 0; i < resp.CDIDevices.size(); i++) { if (i) o += ','; jstr(o, resp.CDIDevices[i]); }
    o += "]}";
    return copy_out(o, json, cap);
}

// ---- snapshot validation of Allocate: tests drive the generation through the seam
void kxh_snapshot_enable(void *h, const uint64_t *generation, const int *healthy) {
    Plugin *p = (Plugin *)h;
    p->snapshotValidation = true;
    if (generation) p->bindGeneration = [generation, healthy](uint64_t &g) { g = *generation; return !healthy || *healthy != 0; };
}
void kxh_validation_counts(void *h, uint64_t *live, uint64_t *snapshot) {
    Plugin *p = (Plugin *)h;
    *live = p->liveValidations;
    *snapshot = p->snapshotValidations;
}
// BindWatcher's uevent parser, message by message (CPU tests): returns the generation after the message
uint64_t kxh_uevent_feed(void **w, const char *msg, size_t len) {
    if (!*w) *w = new device_plugin::BindWatcher();
    device_plugin::BindWatcher *bw = (device_plugin::BindWatcher *)*w;
    if (msg) bw->feed(msg, len);
    return bw->generation();
}
void kxh_uevent_free(void *w) { delete (device_plugin::BindWatcher *)w; }
// the real socket: 0 = opened (and healthy), < 0 = this box does not allow it
int kxh_uevent_socket_ok() {
    device_plugin::BindWatcher bw;
    return bw.start() ? -1 : (bw.healthy() ? 0 : -2);
}

// ---- health watcher (tests): a plugin can be added by hand so that no GPU is needed for the host logic
int kxh_add_plugin(void *h, const char *name, const char *device_path, const char *ids_csv) {
    Plugin *p = (Plugin *)h;
    device_plugin::GenericDevicePlugin dp;
    dp.devpluginName = name;
    dp.devicePath = device_path;
    dp.socketPath = std::string(device_plugin::kDevicePluginPath) + "kata-xpu-" + name + ".sock";
    std::string cur;
    for (const char *c = ids_csv;; c++) {
        if (*c == ',' || *c == 0) {
            if (!cur.empty()) dp.devs.push_back(device_plugin::Device{cur, device_plugin::kHealthy});
            cur.clear();
            if (*c == 0) break;
        } else {
            cur += *c;
        }
    }
    p->devicePlugins.push_back(std::move(dp));
    return (int)p->devicePlugins.size() - 1;
}

void *kxh_health_start(void *h, int plugin_index, int watch_creates, char *err, size_t errcap) {
    Plugin *p = (Plugin *)h;
    if (plugin_index < 0 || (size_t)plugin_index >= p->devicePlugins.size()) return nullptr;
    auto *w = new device_plugin::HealthWatcher(p->devicePlugins[(size_t)plugin_index], watch_creates != 0);
    device_plugin::Error e = w->start();
    if (e) { copy_out(e.message, err, errcap); delete w; return nullptr; }
    return w;
}
int kxh_health_poll(void *w, int timeout_ms) { return ((device_plugin::HealthWatcher *)w)->poll(timeout_ms); }
void kxh_health_stop(void *w) { delete (device_plugin::HealthWatcher *)w; }

// "id=Health,id=Health,..." of one plugin
int kxh_devs(void *h, int plugin_index, char *out, size_t cap) {
    Plugin *p = (Plugin *)h;
    if (plugin_index < 0 || (size_t)plugin_index >= p->devicePlugins.size()) return -1;
    std::string o;
    for (const auto &d : p->devicePlugins[(size_t)plugin_index].devs) {
        if (!o.empty()) o += ',';
        o += d.ID + "=" + d.Health;
    }
    return copy_out(o, out, cap);
}

int kxh_list_and_watch(void *h, int plugin_index, uint8_t *out, size_t cap) {
    Plugin *p = (Plugin *)h;
    if (plugin_index < 0 || (size_t)plugin_index >= p->devicePlugins.size()) return -1;
    std::vector<uint8_t> b;
    if (p->ListAndWatchBytes(p->devicePlugins[(size_t)plugin_index], b)) return -1;
    if (b.size() > cap) return -2;
    memcpy(out, b.data(), b.size());
    return (int)b.size();
}

}  // extern "C"
