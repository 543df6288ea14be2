// comm.cu -- multi-GPU pci.ids load: one rank per GPU, shards cut at vendor-line boundaries,
// ONE ncclAllGather of hit rows over NVLink, then a min-merge on every rank.
//
// BASELINE.json configs[3] / SURVEY.md 8(e): every rank parses its byte range of one logical
// text (offsets are GLOBAL: global_base + local), which yields for every (vendor,device) key
// the earliest candidate line of the shard, and for every vendor prefix its earliest anchor.
// Candidates of all ranks are exchanged as fixed-capacity slabs
//     [header | key rows | vendor rows | sanitised names]
// and folded into a fresh table with the same atomicMin rule the parse kernel uses, so the
// result equals kxpu_pciids_load on the concatenated text.  The payload is ~1 MB per rank
// (18 856 rows x 32 B + 0.8 MB of names for pci.ids), i.e. latency bound on NVSwitch.
//
// NCCL is loaded lazily with dlopen: a single-GPU deployment never needs libnccl.
#include <dlfcn.h>

#include <algorithm>
#include <new>

#include "common.cuh"
#include "slab.cuh"
#include "table.cuh"

// from api.cu
struct kxpu_table;
int32_t kx_build_table(kxpu_ctx *ctx, const uint8_t *d_text, size_t n, unsigned long long base,
                       unsigned long long carry_in, int check_valid, kxpu_table **out);
int32_t kx_table_from_gather(kxpu_ctx *ctx, void *d_gather, int nranks, size_t slab_stride, kxcomm::SlabCaps caps,
                             kxpu_table **out, bool own_names, const uint32_t *peer_timeout);
void kx_table_local_view(kxpu_table *t, KxTableDev *dev, uint32_t *cap, uint32_t *n_rows, uint32_t *blob_used,
                         const uint32_t **row_key, const unsigned long long **row_line, const unsigned long long **row_anchor,
                         const uint32_t **row_name_off, const uint32_t **row_name_len, const uint8_t **blob);
void kx_table_release(kxpu_ctx *ctx, kxpu_table *t);

namespace kxcomm {

typedef int (*fn_get_unique_id)(void *);
typedef int (*fn_comm_init_rank)(void **, int, char[128], int);  // ncclUniqueId is passed by value (128 bytes)
typedef int (*fn_comm_destroy)(void *);
typedef int (*fn_all_gather)(const void *, void *, size_t, int, void *, cudaStream_t);
typedef const char *(*fn_err_string)(int);

struct NcclApi {
    void *handle = nullptr;
    fn_get_unique_id get_unique_id = nullptr;
    void *comm_init_rank = nullptr;
    fn_comm_destroy comm_destroy = nullptr;
    fn_all_gather all_gather = nullptr;
    fn_err_string err_string = nullptr;
};

static NcclApi g_nccl;
static std::mutex g_nccl_mu;

static bool nccl_load() {
    std::lock_guard<std::mutex> g(g_nccl_mu);
    if (g_nccl.handle) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *nm : names) {
        h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) return false;
    g_nccl.get_unique_id = (fn_get_unique_id)dlsym(h, "ncclGetUniqueId");
    g_nccl.comm_init_rank = dlsym(h, "ncclCommInitRank");
    g_nccl.comm_destroy = (fn_comm_destroy)dlsym(h, "ncclCommDestroy");
    g_nccl.all_gather = (fn_all_gather)dlsym(h, "ncclAllGather");
    g_nccl.err_string = (fn_err_string)dlsym(h, "ncclGetErrorString");
    if (!g_nccl.get_unique_id || !g_nccl.comm_init_rank || !g_nccl.comm_destroy || !g_nccl.all_gather) return false;
    g_nccl.handle = h;
    return true;
}

struct UniqueId { char internal[128]; };
typedef int (*fn_comm_init_rank_byval)(void **, int, UniqueId, int);

__global__ void __launch_bounds__(256) pack_rows_kernel(uint8_t *slab, SlabCaps caps, uint32_t n_rows,
                                                        const uint32_t *row_key, const unsigned long long *row_line,
                                                        const unsigned long long *row_anchor, const uint32_t *row_name_off,
                                                        const uint32_t *row_name_len, const unsigned long long *trunc,
                                                        uint32_t blob_used) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    SlabHeader *h = reinterpret_cast<SlabHeader *>(slab);
    if (i == 0) {
        h->n_rows = n_rows < caps.rows ? n_rows : caps.rows;
        h->blob_bytes = blob_used;
        h->trunc = *trunc;
        h->reserved = 0;
        if (n_rows > caps.rows || blob_used > caps.blob) atomicOr(&h->overflow, 1u);
    }
    if (i >= n_rows || i >= caps.rows) return;
    SlabRow *rows = reinterpret_cast<SlabRow *>(slab + sizeof(SlabHeader));
    SlabRow r;
    r.key = row_key[i]; r.name_len = row_name_len[i]; r.line = row_line[i]; r.anchor = row_anchor[i];
    r.name_off = row_name_off[i]; r.pad = 0;
    rows[i] = r;
}

__global__ void __launch_bounds__(256) pack_vendors_kernel(uint8_t *slab, SlabCaps caps, const unsigned long long *vendor_first) {
    uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;  // 65536 threads
    SlabHeader *h = reinterpret_cast<SlabHeader *>(slab);
    unsigned long long f = vendor_first[v];
    if (f == KX_NO_OFF) return;
    uint32_t idx = atomicAdd(&h->n_vendors, 1u);
    if (idx >= caps.vendors) { atomicOr(&h->overflow, 2u); return; }
    SlabVendor *vs = reinterpret_cast<SlabVendor *>(slab + sizeof(SlabHeader) + (size_t)caps.rows * sizeof(SlabRow));
    vs[idx].vendor = v; vs[idx].pad = 0; vs[idx].first = f;
}

}  // namespace kxcomm

#include "p2p.cuh"

using namespace kxcomm;

extern "C" int32_t kxpu_comm_unique_id(uint8_t id_out[KXPU_COMM_ID_BYTES]) {
    if (!id_out) return KXPU_E_INVALID;
    if (!nccl_load()) return KXPU_E_NCCL;
    return g_nccl.get_unique_id(id_out) == 0 ? KXPU_OK : KXPU_E_NCCL;
}

extern "C" int32_t kxpu_comm_init(kxpu_ctx *ctx, int32_t nranks, int32_t rank, const uint8_t id[KXPU_COMM_ID_BYTES]) {
    if (!ctx || !id || nranks < 1 || rank < 0 || rank >= nranks) return KXPU_E_INVALID;
    std::lock_guard<std::mutex> guard(ctx->mu);
    cudaSetDevice(ctx->device);
    if (!nccl_load()) { KX_SET_ERR(ctx, "libnccl.so.2 not found"); return KXPU_E_NCCL; }
    if (ctx->nccl_comm) return KXPU_E_INVALID;
    UniqueId uid;
    memcpy(uid.internal, id, 128);
    void *comm = nullptr;
    int rc = ((fn_comm_init_rank_byval)g_nccl.comm_init_rank)(&comm, nranks, uid, rank);
    if (rc != 0) {
        KX_SET_ERR(ctx, "ncclCommInitRank: %s", g_nccl.err_string ? g_nccl.err_string(rc) : "error");
        return KXPU_E_NCCL;
    }
    ctx->nccl_comm = comm;
    ctx->nranks = nranks;
    ctx->rank = rank;
    p2p_setup(ctx, [&](const void *src, void *dst, size_t nbytes) {
        return g_nccl.all_gather(src, dst, nbytes, /*ncclUint8*/ 1, ctx->nccl_comm, ctx->stream);
    });
    return KXPU_OK;
}

extern "C" int32_t kxpu_comm_destroy(kxpu_ctx *ctx) {
    if (!ctx) return KXPU_E_INVALID;
    std::lock_guard<std::mutex> guard(ctx->mu);
    cudaSetDevice(ctx->device);
    if (ctx->nccl_comm) {
        cudaStreamSynchronize(ctx->stream);
        p2p_teardown(ctx);
        g_nccl.comm_destroy(ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
        ctx->nranks = 1;
        ctx->rank = 0;
    }
    return KXPU_OK;
}

extern "C" int32_t kxpu_pciids_load_sharded(kxpu_ctx *ctx, const void *d_text_shard, size_t n, uint64_t global_base,
                                            kxpu_table **out) {
    if (!ctx || !out || (!d_text_shard && n)) return KXPU_E_INVALID;
    std::lock_guard<std::mutex> guard(ctx->mu);
    cudaSetDevice(ctx->device);
    kx_clear_timings(ctx);
    if (!ctx->nccl_comm) { KX_SET_ERR(ctx, "kxpu_comm_init has not been called"); return KXPU_E_NCCL; }
    const int R = ctx->nranks;

    // 1. local parse: every candidate row of the shard (validity is a global property)
    kxpu_table *local = nullptr;
    int32_t rc = kx_build_table(ctx, (const uint8_t *)d_text_shard, n, global_base, 0, 0, &local);
    if (rc != KXPU_OK) return rc;
    KxTableDev dev;
    uint32_t cap, n_rows, blob_used;
    const uint32_t *row_key, *row_name_off, *row_name_len;
    const unsigned long long *row_line, *row_anchor;
    const uint8_t *blob;
    kx_table_local_view(local, &dev, &cap, &n_rows, &blob_used, &row_key, &row_line, &row_anchor, &row_name_off,
                        &row_name_len, &blob);

    // ---- peer-memory path: pack into my slot, push it to every peer, wait for everybody's flag, merge
    if (ctx->p2p_ok) {
        const SlabCaps caps = kDefaultCaps;
        const uint32_t e = ++ctx->p2p_epoch;
        const int b = (int)(e & 1u);
        const size_t slot0 = P2P_FLAGS_BYTES + (size_t)b * (size_t)R * ctx->p2p_stride;
        if (ctx->stage_timing) cudaEventRecord(ctx->ev[2 * KXPU_T_MERGE], ctx->stream);
        static const bool ptrace = getenv("KXPU_TRACE_MERGE") != nullptr;
        static cudaEvent_t pev[4];
        static bool pev_ok = false;
        if (ptrace && !pev_ok) { for (auto &e2 : pev) cudaEventCreate(&e2); pev_ok = true; }
        if (ptrace) cudaEventRecord(pev[0], ctx->stream);
        PushParams PP;
        memset(&PP, 0, sizeof PP);
        PP.nranks = R; PP.epoch = e; PP.caps = caps; PP.scratch = ctx->p2p_scratch;
        PP.n_rows = n_rows; PP.blob_used = blob_used; PP.row_key = row_key; PP.row_name_off = row_name_off;
        PP.row_name_len = row_name_len; PP.row_line = row_line; PP.row_anchor = row_anchor; PP.trunc = dev.trunc;
        PP.vendor_first = dev.vendor_first; PP.blob = blob;
        for (int q = 0; q < R; q++) {
            PP.dst[q] = ctx->p2p_peer[q] + slot0 + (size_t)ctx->rank * ctx->p2p_stride;
            PP.flag[q] = reinterpret_cast<uint32_t *>(ctx->p2p_peer[q]) + b * kxpu_ctx::KX_P2P_MAX_RANKS + ctx->rank;
        }
        pack_push_kernel<<<2 * ctx->sm_count, 256, 0, ctx->stream>>>(PP);
        if (ptrace) cudaEventRecord(pev[1], ctx->stream);
        wait_flags_kernel<<<1, 32, 0, ctx->stream>>>(reinterpret_cast<const uint32_t *>(ctx->p2p_local) + b * kxpu_ctx::KX_P2P_MAX_RANKS,
                                                     R, e, ctx->p2p_scratch + 1);
        ctx->launches += 2;
        if (ptrace) cudaEventRecord(pev[2], ctx->stream);
        kxpu_table *merged = nullptr;
        rc = kx_table_from_gather(ctx, ctx->p2p_local + slot0, R, ctx->p2p_stride, caps, &merged, /*own_names=*/true, ctx->p2p_scratch + 1);
        if (ptrace) {
            cudaEventRecord(pev[3], ctx->stream);
            cudaEventSynchronize(pev[3]);
            float a = 0, b2 = 0, c2 = 0;
            cudaEventElapsedTime(&a, pev[0], pev[1]); cudaEventElapsedTime(&b2, pev[1], pev[2]); cudaEventElapsedTime(&c2, pev[2], pev[3]);
            static int calls = 0;
            if (++calls % 8 == 0) fprintf(stderr, "[kxpu merge trace] pack+push %.1f us  wait %.1f us  merge+sync %.1f us  (peer memory, %d ranks)\n", a * 1e3, b2 * 1e3, c2 * 1e3, R);
        }
        if (ctx->stage_timing) { cudaEventRecord(ctx->ev[2 * KXPU_T_MERGE + 1], ctx->stream); ctx->ev_used[KXPU_T_MERGE] = true; }
        if (rc == KXPU_E_NCCL) {  // a peer never delivered (flagged by the wait kernel, seen by the merge)
            cudaMemsetAsync(ctx->p2p_scratch, 0, 256, ctx->stream);
            kx_table_release(ctx, local);
            return rc;
        }
        if (rc != KXPU_E_CAPACITY) {  // a slab outgrew the fixed exchange region: every rank falls back to NCCL below
            kx_table_release(ctx, local);
            if (rc != KXPU_OK) return rc;
            *out = merged;
            return KXPU_OK;
        }
    }

    SlabCaps caps{32768u, 8192u, 1u << 20};
    if (ctx->p2p_ok) { caps.rows *= 4; caps.vendors = 65536; caps.blob *= 8; }  // the default capacities just overflowed
    for (int attempt = 0; attempt < 6; attempt++) {
        const size_t sb = (slab_bytes(caps) + 255) / 256 * 256;
        uint8_t *d_slab = nullptr, *d_gather = nullptr;
        cudaError_t e = cudaMallocAsync((void **)&d_slab, sb, ctx->stream);
        if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_gather, sb * (size_t)R, ctx->stream);
        if (e != cudaSuccess) {
            KX_SET_ERR(ctx, "gather buffers: %s", cudaGetErrorString(e));
            kx_table_release(ctx, local);
            return KXPU_E_NOMEM;
        }
        if (ctx->stage_timing) cudaEventRecord(ctx->ev[2 * KXPU_T_MERGE], ctx->stream);
        static const bool trace = getenv("KXPU_TRACE_MERGE") != nullptr;
        static cudaEvent_t tev[4];
        static bool tev_ok = false;
        if (trace && !tev_ok) { for (auto &e2 : tev) cudaEventCreate(&e2); tev_ok = true; }
        if (trace) cudaEventRecord(tev[0], ctx->stream);
        // 2. pack the slab
        cudaMemsetAsync(d_slab, 0, sizeof(SlabHeader), ctx->stream);
        pack_rows_kernel<<<(std::max(n_rows, 1u) + 255) / 256, 256, 0, ctx->stream>>>(
            d_slab, caps, n_rows, row_key, row_line, row_anchor, row_name_off, row_name_len, dev.trunc, blob_used);
        pack_vendors_kernel<<<65536 / 256, 256, 0, ctx->stream>>>(d_slab, caps, dev.vendor_first);
        ctx->launches += 2;
        if (blob_used > 0 && blob_used <= caps.blob)
            cudaMemcpyAsync(d_slab + slab_blob_off(caps), blob, blob_used, cudaMemcpyDeviceToDevice, ctx->stream);
        if (trace) cudaEventRecord(tev[1], ctx->stream);
        // 3. the one collective of the path
        int nrc = g_nccl.all_gather(d_slab, d_gather, sb, /*ncclUint8*/ 1, ctx->nccl_comm, ctx->stream);
        if (nrc != 0) {
            KX_SET_ERR(ctx, "ncclAllGather: %s", g_nccl.err_string ? g_nccl.err_string(nrc) : "error");
            cudaFreeAsync(d_slab, ctx->stream);
            cudaFreeAsync(d_gather, ctx->stream);
            kx_table_release(ctx, local);
            return KXPU_E_NCCL;
        }
        if (trace) cudaEventRecord(tev[2], ctx->stream);
        cudaFreeAsync(d_slab, ctx->stream);
        // 4. min-merge into a fresh table (keeps d_gather: the names live there)
        kxpu_table *merged = nullptr;
        rc = kx_table_from_gather(ctx, d_gather, R, sb, caps, &merged, /*own_names=*/false, nullptr);
        if (ctx->stage_timing) { cudaEventRecord(ctx->ev[2 * KXPU_T_MERGE + 1], ctx->stream); ctx->ev_used[KXPU_T_MERGE] = true; }
        if (trace) {
            cudaEventRecord(tev[3], ctx->stream);
            cudaEventSynchronize(tev[3]);
            float a = 0, b = 0, c2 = 0;
            cudaEventElapsedTime(&a, tev[0], tev[1]); cudaEventElapsedTime(&b, tev[1], tev[2]); cudaEventElapsedTime(&c2, tev[2], tev[3]);
            static int calls = 0;
            if (++calls % 8 == 0) fprintf(stderr, "[kxpu merge trace] pack %.1f us  all-gather %.1f us  merge+sync %.1f us  slab %zu B x %d\n", a * 1e3, b * 1e3, c2 * 1e3, sb, R);
        }
        if (rc == KXPU_E_CAPACITY) {
            // some rank overflowed a slab capacity: every rank sees the same headers and retries alike
            caps.rows *= 4; caps.vendors = 65536; caps.blob *= 8;
            continue;
        }
        kx_table_release(ctx, local);
        if (rc != KXPU_OK) return rc;
        *out = merged;
        return KXPU_OK;
    }
    kx_table_release(ctx, local);
    return KXPU_E_CAPACITY;
}
