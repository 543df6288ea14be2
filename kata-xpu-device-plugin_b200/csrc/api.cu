// api.cu -- context, memory helpers and the pci.ids entry points of the C ABI (include/kxpu.h).
#include <algorithm>
#include <new>
#include <vector>

#include "pciids.cu"  // kernels (single translation unit keeps them inlinable and static)
#include "pciids2.cu"
#include "pciids3.cu"
#include "pciids4.cu"
#include "pciids5.cu"
#include "slab.cuh"

#include <cstdlib>

struct kxpu_table {
    uint32_t cap = 0, shift = 0;
    void *arena = nullptr;
    size_t arena_bytes = 0;
    KxTableDev dev{};
    unsigned long long *tile_state = nullptr;
    int32_t *row_of_slot = nullptr;
    uint32_t *row_key = nullptr;
    unsigned long long *row_line = nullptr;
    unsigned long long *row_anchor = nullptr;
    uint32_t *row_name_off = nullptr;
    uint32_t *row_name_len = nullptr;
    uint32_t *sel = nullptr;
    uint8_t *blob = nullptr;
    uint32_t blob_cap = 0;
    uint32_t n_rows = 0;
    uint32_t blob_used = 0;
    void *gather = nullptr;  // sharded load: all-gathered row slabs (names live here)
};

static const char *kx_err_names[] = {
    "ok", "invalid argument", "CUDA error", "no sm_100 GPU available", "output buffer too small",
    "table capacity exceeded", "NCCL unavailable or failed", "input outside the supported domain", "out of memory"};

extern "C" const char *kxpu_strerror(int32_t status) {
    int i = -status;
    if (i < 0 || i > 8) return "unknown status";
    return kx_err_names[i];
}

extern "C" const char *kxpu_last_error(kxpu_ctx *ctx) { return ctx ? ctx->err : "null ctx"; }
extern "C" uint64_t kxpu_launch_count(kxpu_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int32_t kxpu_ctx_create(int32_t ordinal, kxpu_ctx **out) {
    if (!out) return KXPU_E_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || ordinal < 0 || ordinal >= count) {
        cudaGetLastError();
        return KXPU_E_NOGPU;  // no CPU fallback: the caller must treat this as fatal
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, ordinal) != cudaSuccess) return KXPU_E_NOGPU;
    if (prop.major != 10) return KXPU_E_NOGPU;  // kernels are built for sm_100a only
    kxpu_ctx *c = new (std::nothrow) kxpu_ctx();
    if (!c) return KXPU_E_NOMEM;
    c->device = ordinal;
    c->sm_count = prop.multiProcessorCount;
    if (cudaSetDevice(ordinal) != cudaSuccess || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete c;
        return KXPU_E_CUDA;
    }
    for (int i = 0; i < 2 * KXPU_T_COUNT; i++) cudaEventCreate(&c->ev[i]);
    cudaEventCreate(&c->ev_user[0]);
    cudaEventCreate(&c->ev_user[1]);
    cudaMallocHost((void **)&c->h_ctl, 64 * sizeof(uint32_t));
    // keep stream-ordered allocations cached: table builds allocate/free per call
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, ordinal) == cudaSuccess) {
        uint64_t thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    cudaFuncSetAttribute(kxparse::parse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)sizeof(kxparse::ParseSmem));
    cudaFuncSetAttribute(kxparse2::parse_kernel_v2, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)(sizeof(kxparse2::WarpSmem) * kxparse2::WARPS));
    cudaFuncSetAttribute(kxparse3::parse_kernel_v3, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)sizeof(kxparse3::CtaSmem3));
    // 1 = CTA-tiled kernel (pciids.cu), 2 = warp-autonomous (pciids2.cu), 3 = super-chunk (pciids3.cu, default)
    cudaFuncSetAttribute(kxparse4::parse_kernel_v4, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)sizeof(kxparse4::CtaSmem4));
    const char *pv = getenv("KXPU_PARSE_V");
    cudaFuncSetAttribute(kxparse5::parse_kernel_v5, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)(sizeof(kxparse5::WarpSmem5) * kxparse2::WARPS));
    c->parse_version = (pv && pv[0] >= '1' && pv[0] <= '5') ? pv[0] - '0' : 5;
    const char *fr = getenv("KXPU_RCH");
    c->force_rch = (fr && fr[0] >= '1' && fr[0] <= '8' && !fr[1]) ? fr[0] - '0' : 0;
    *out = c;
    return KXPU_OK;
}

extern "C" int32_t kxpu_ctx_destroy(kxpu_ctx *ctx) {
    if (!ctx) return KXPU_E_INVALID;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < 2 * KXPU_T_COUNT; i++) cudaEventDestroy(ctx->ev[i]);
    cudaEventDestroy(ctx->ev_user[0]);
    cudaEventDestroy(ctx->ev_user[1]);
    cudaFreeHost(ctx->h_ctl);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return KXPU_OK;
}

extern "C" int32_t kxpu_set_stage_timing(kxpu_ctx *ctx, int32_t on) {
    if (!ctx) return KXPU_E_INVALID;
    std::lock_guard<std::mutex> g(ctx->mu);
    ctx->stage_timing = on != 0;
    return KXPU_OK;
}

extern "C" int32_t kxpu_last_timings(kxpu_ctx *ctx, float ms_out[KXPU_T_COUNT]) {
    if (!ctx || !ms_out) return KXPU_E_INVALID;
    std::lock_guard<std::mutex> g(ctx->mu);
    cudaSetDevice(ctx->device);
    KX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < KXPU_T_COUNT; i++) {
        ms_out[i] = 0.f;
        if (ctx->ev_used[i]) cudaEventElapsedTime(&ms_out[i], ctx->ev[2 * i], ctx->ev[2 * i + 1]);
    }
    return KXPU_OK;
}

// ------------------------------------------------------------------ memory helpers
#define KX_ENTER(ctx)                          \
    if (!(ctx)) return KXPU_E_INVALID;         \
    std::lock_guard<std::mutex> guard__((ctx)->mu); \
    cudaSetDevice((ctx)->device)

extern "C" int32_t kxpu_timer_begin(kxpu_ctx *ctx) {
    KX_ENTER(ctx);
    KX_CUDA(ctx, cudaEventRecord(ctx->ev_user[0], ctx->stream));
    return KXPU_OK;
}
extern "C" int32_t kxpu_timer_end(kxpu_ctx *ctx, float *ms_out) {
    KX_ENTER(ctx);
    if (!ms_out) return KXPU_E_INVALID;
    KX_CUDA(ctx, cudaEventRecord(ctx->ev_user[1], ctx->stream));
    KX_CUDA(ctx, cudaEventSynchronize(ctx->ev_user[1]));
    KX_CUDA(ctx, cudaEventElapsedTime(ms_out, ctx->ev_user[0], ctx->ev_user[1]));
    return KXPU_OK;
}

extern "C" int32_t kxpu_dev_alloc(kxpu_ctx *ctx, size_t bytes, void **d_out) {
    KX_ENTER(ctx);
    if (!d_out) return KXPU_E_INVALID;
    KX_CUDA(ctx, cudaMalloc(d_out, bytes ? bytes : 16));
    return KXPU_OK;
}
extern "C" int32_t kxpu_dev_free(kxpu_ctx *ctx, void *d_ptr) {
    KX_ENTER(ctx);
    KX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    KX_CUDA(ctx, cudaFree(d_ptr));
    return KXPU_OK;
}
extern "C" int32_t kxpu_dev_upload(kxpu_ctx *ctx, void *d_dst, const void *h_src, size_t bytes) {
    KX_ENTER(ctx);
    KX_CUDA(ctx, cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    KX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KXPU_OK;
}
extern "C" int32_t kxpu_dev_download(kxpu_ctx *ctx, void *h_dst, const void *d_src, size_t bytes) {
    KX_ENTER(ctx);
    KX_CUDA(ctx, cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    KX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KXPU_OK;
}
extern "C" int32_t kxpu_dev_replicate(kxpu_ctx *ctx, void *d_dst, const void *d_src, size_t n, size_t copies) {
    KX_ENTER(ctx);
    // doubling copy: log2(copies) device-to-device memcpys
    if (copies == 0 || n == 0) return KXPU_OK;
    uint8_t *dst = (uint8_t *)d_dst;
    if (dst != d_src) KX_CUDA(ctx, cudaMemcpyAsync(dst, d_src, n, cudaMemcpyDeviceToDevice, ctx->stream));
    size_t have = 1;
    while (have < copies) {
        size_t add = std::min(have, copies - have);
        KX_CUDA(ctx, cudaMemcpyAsync(dst + have * n, dst, add * n, cudaMemcpyDeviceToDevice, ctx->stream));
        have += add;
    }
    KX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KXPU_OK;
}
extern "C" int32_t kxpu_pinned_alloc(kxpu_ctx *ctx, size_t bytes, void **h_out) {
    KX_ENTER(ctx);
    if (!h_out) return KXPU_E_INVALID;
    KX_CUDA(ctx, cudaMallocHost(h_out, bytes ? bytes : 16));
    return KXPU_OK;
}
extern "C" int32_t kxpu_pinned_free(kxpu_ctx *ctx, void *h_ptr) {
    KX_ENTER(ctx);
    KX_CUDA(ctx, cudaFreeHost(h_ptr));
    return KXPU_OK;
}
extern "C" int32_t kxpu_sync(kxpu_ctx *ctx) {
    KX_ENTER(ctx);
    KX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KXPU_OK;
}

// ------------------------------------------------------------------ table build
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static void table_release(kxpu_ctx *ctx, kxpu_table *t) {
    if (!t) return;
    if (t->arena) cudaFreeAsync(t->arena, ctx->stream);
    if (t->gather) cudaFreeAsync(t->gather, ctx->stream);
    delete t;
}

static int32_t table_alloc(kxpu_ctx *ctx, uint32_t cap, uint32_t blob_cap, uint32_t num_tiles, kxpu_table **out,
                           bool zero_tiles = true) {
    kxpu_table *t = new (std::nothrow) kxpu_table();
    if (!t) return KXPU_E_NOMEM;
    t->cap = cap;
    uint32_t lg = 0;
    while ((1u << lg) < cap) lg++;
    t->shift = 32 - lg;
    const size_t slots = (size_t)cap + 1;
    size_t off = 0;
    // 0xff-initialised region first (one memset), then the zero-initialised region
    size_t o_keys = off;        off = align_up(off + slots * 4, 256);
    size_t o_min_line = off;    off = align_up(off + slots * 8, 256);
    size_t o_min_anchor = off;  off = align_up(off + slots * 8, 256);
    size_t o_vfirst = off;      off = align_up(off + 65536 * 8, 256);
    size_t o_trunc = off;       off = align_up(off + 8, 256);
    size_t ff_bytes = off;
    size_t o_counters = off;    off = align_up(off + KX_C_COUNT * 4, 256);
    // v1-v4: one zeroed status word per tile; v5: up to 24 B of (never zeroed) range words per chunk
    size_t o_tiles = off;       off = align_up(off + ((size_t)num_tiles + 64) * (zero_tiles ? 8 : 24), 256);
    size_t zero_bytes = off - ff_bytes;
    size_t o_row_of_slot = off; off = align_up(off + slots * 4, 256);
    size_t o_row_key = off;     off = align_up(off + slots * 4, 256);
    size_t o_row_line = off;    off = align_up(off + slots * 8, 256);
    size_t o_row_anchor = off;  off = align_up(off + slots * 8, 256);
    size_t o_row_noff = off;    off = align_up(off + slots * 4, 256);
    size_t o_row_nlen = off;    off = align_up(off + slots * 4, 256);
    size_t o_sel = off;         off = align_up(off + slots * 4, 256);
    size_t o_blob = off;        off = align_up(off + blob_cap, 256);
    t->arena_bytes = off;
    cudaError_t e = cudaMallocAsync(&t->arena, off, ctx->stream);
    if (e != cudaSuccess) {
        KX_SET_ERR(ctx, "cudaMallocAsync(%zu) -> %s", off, cudaGetErrorString(e));
        delete t;
        return e == cudaErrorMemoryAllocation ? KXPU_E_NOMEM : KXPU_E_CUDA;
    }
    uint8_t *b = (uint8_t *)t->arena;
    t->dev.keys = (uint32_t *)(b + o_keys);
    t->dev.min_line = (unsigned long long *)(b + o_min_line);
    t->dev.min_anchor = (unsigned long long *)(b + o_min_anchor);
    t->dev.vendor_first = (unsigned long long *)(b + o_vfirst);
    t->dev.trunc = (unsigned long long *)(b + o_trunc);
    t->dev.counters = (uint32_t *)(b + o_counters);
    t->dev.cap = cap;
    t->dev.shift = t->shift;
    t->dev.max_keys = cap / 2;
    t->tile_state = (unsigned long long *)(b + o_tiles);
    t->row_of_slot = (int32_t *)(b + o_row_of_slot);
    t->row_key = (uint32_t *)(b + o_row_key);
    t->row_line = (unsigned long long *)(b + o_row_line);
    t->row_anchor = (unsigned long long *)(b + o_row_anchor);
    t->row_name_off = (uint32_t *)(b + o_row_noff);
    t->row_name_len = (uint32_t *)(b + o_row_nlen);
    t->sel = (uint32_t *)(b + o_sel);
    t->blob = b + o_blob;
    t->blob_cap = blob_cap;
    cudaMemsetAsync(b, 0xff, ff_bytes, ctx->stream);
    // the v5 parse kernels write every status word they read: only the counters need zeroing
    cudaMemsetAsync(b + ff_bytes, 0, zero_tiles ? zero_bytes : o_tiles - ff_bytes, ctx->stream);
    *out = t;
    return KXPU_OK;
}

static int parse_grid(kxpu_ctx *ctx, uint32_t num_tiles) {
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kxparse::parse_kernel, kxparse::NT,
                                                  sizeof(kxparse::ParseSmem));
    if (per_sm < 1) per_sm = 1;
    long long g = (long long)per_sm * ctx->sm_count;
    if (g > (long long)num_tiles) g = num_tiles;
    return (int)(g < 1 ? 1 : g);
}

static int parse_grid_v2(kxpu_ctx *ctx, uint32_t num_chunks) {
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kxparse2::parse_kernel_v2, kxparse2::NT,
                                                  sizeof(kxparse2::WarpSmem) * kxparse2::WARPS);
    if (per_sm < 1) per_sm = 1;
    long long g = (long long)per_sm * ctx->sm_count;
    long long need = ((long long)num_chunks + kxparse2::WARPS - 1) / kxparse2::WARPS;
    if (g > need) g = need;
    return (int)(g < 1 ? 1 : g);
}

static int parse_grid_v3(kxpu_ctx *ctx, uint32_t num_sc) {
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kxparse3::parse_kernel_v3, kxparse2::NT, sizeof(kxparse3::CtaSmem3));
    if (per_sm < 1) per_sm = 1;
    long long g = (long long)per_sm * ctx->sm_count;
    if (g > (long long)num_sc) g = num_sc;
    return (int)(g < 1 ? 1 : g);
}

static int parse_grid_v4(kxpu_ctx *ctx, uint32_t num_ranges) {
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kxparse4::parse_kernel_v4, kxparse2::NT, sizeof(kxparse4::CtaSmem4));
    if (per_sm < 1) per_sm = 1;
    long long g = (long long)per_sm * ctx->sm_count;
    if (g > (long long)num_ranges) g = num_ranges;
    return (int)(g < 1 ? 1 : g);
}

static int32_t launch_finalize(kxpu_ctx *ctx, kxpu_table *t, const uint8_t *d_text, size_t n,
                               unsigned long long base, int check_valid) {
    kxparse::FinalizeParams F;
    F.text = d_text; F.n = n; F.base = base; F.tab = t->dev;
    F.row_of_slot = t->row_of_slot; F.row_key = t->row_key; F.row_line = t->row_line; F.row_anchor = t->row_anchor;
    F.row_name_off = t->row_name_off; F.row_name_len = t->row_name_len; F.sel = t->sel;
    F.blob = t->blob; F.blob_cap = t->blob_cap; F.check_valid = check_valid;
    kxparse::finalize_select_kernel<<<(t->cap + 1 + 255) / 256, 256, 0, ctx->stream>>>(F);
    // one warp per selected slot; warps beyond the (device-side) count exit at once
    kxparse::finalize_kernel<<<(t->cap + 1 + kxparse::FIN_WARPS - 1) / kxparse::FIN_WARPS, kxparse::FIN_WARPS * 32, 0, ctx->stream>>>(F);
    ctx->launches += 2;
    KX_CUDA(ctx, cudaGetLastError());
    return KXPU_OK;
}

struct KxJoin {  // optional batched join enqueued behind the finalize, in front of the host round trip
    const uint32_t *d_keys;
    size_t n;
    int32_t *d_rows;
};
static int32_t launch_lookup(kxpu_ctx *ctx, kxpu_table *t, const uint32_t *d_keys, size_t n, int32_t *d_rows);

// Parse d_text[0..n) (global offsets base..base+n) and finalize.  check_valid=1 yields the
// final table of a single text; 0 leaves every local candidate row for the sharded merge.
static int32_t kx_build_table_join(kxpu_ctx *ctx, const uint8_t *d_text, size_t n, unsigned long long base,
                                   unsigned long long carry_in, int check_valid, kxpu_table **out, const KxJoin *join);
int32_t kx_build_table(kxpu_ctx *ctx, const uint8_t *d_text, size_t n, unsigned long long base,
                       unsigned long long carry_in, int check_valid, kxpu_table **out) {
    return kx_build_table_join(ctx, d_text, n, base, carry_in, check_valid, out, nullptr);
}
static int32_t kx_build_table_join(kxpu_ctx *ctx, const uint8_t *d_text, size_t n, unsigned long long base,
                                   unsigned long long carry_in, int check_valid, kxpu_table **out, const KxJoin *join) {
    if ((reinterpret_cast<uintptr_t>(d_text) & 15u) != 0) {
        KX_SET_ERR(ctx, "device text pointer must be 16-byte aligned");
        return KXPU_E_INVALID;
    }
    if (base + n >= (1ull << 44)) return KXPU_E_UNSUPPORTED;
    int version = ctx->parse_version;
    uint32_t cap = 1u << 16;
    uint32_t blob_cap = (uint32_t)std::min<size_t>(std::max<size_t>(n, 256), 4u << 20);
    for (int attempt = 0; attempt < 8; attempt++) {
        kxpu_table *t = nullptr;
        const uint32_t tile_bytes = version == 1 ? (uint32_t)kxparse::T : (uint32_t)kxparse2::CW;
        const uint32_t num_tiles = (uint32_t)((n + tile_bytes - 1) / tile_bytes);
        const uint32_t num_sc = (num_tiles + kxparse3::SCC - 1) / kxparse3::SCC;
        int32_t rc = table_alloc(ctx, cap, blob_cap, num_tiles, &t, version != 5);
        if (rc != KXPU_OK) return rc;
        if (num_tiles > 0) {
            KxTimer tm(ctx, KXPU_T_PARSE);
            if (version == 5) {
                kxparse5::Params5 P;
                P.text = d_text; P.n = n; P.base = base; P.num_chunks = num_tiles;
                P.tma_limit = n >= (size_t)kxparse2::STG_BYTES ? (uint32_t)((n - kxparse2::STG_BYTES) / kxparse2::CW) + 1u : 0u;
                int per_sm = 0;
                const size_t smem = sizeof(kxparse5::WarpSmem5) * kxparse2::WARPS;
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kxparse5::parse_kernel_v5, kxparse2::NT, smem);
                if (per_sm < 1) per_sm = 1;
                // ranges of 8 chunks; shorter ones when the text is too small to give every warp a range
                const uint32_t wave_warps = (uint32_t)per_sm * ctx->sm_count * kxparse2::WARPS;
                P.rch = std::min<uint32_t>(std::max<uint32_t>(num_tiles / wave_warps, 1u), kxparse5::RCH5_MAX);
                if (ctx->force_rch) P.rch = (uint32_t)ctx->force_rch;
                P.num_ranges = (num_tiles + P.rch - 1) / P.rch;
                P.range_state = t->tile_state;                            // [num_ranges]
                P.range_carry = t->tile_state + P.num_ranges;             // [num_ranges]
                P.lead = (uint32_t *)(t->tile_state + 2 * P.num_ranges);  // [num_ranges]
                P.tasks = P.lead + P.num_ranges;                          // [num_tiles]
                P.tab = t->dev; P.carry_in = carry_in;
                // persistent grid, at most one warp per range
                uint32_t grid = (uint32_t)per_sm * ctx->sm_count;
                const uint32_t need = (P.num_ranges + kxparse2::WARPS - 1) / kxparse2::WARPS;
                if (grid > need) grid = need;
                kxparse5::parse_kernel_v5<<<grid, kxparse2::NT, smem, ctx->stream>>>(P);
                KX_LAUNCHED(ctx);
                tm.stop();
                KxTimer tr(ctx, KXPU_T_RESOLVE);
                kxparse5::resolve_ranges_kernel<<<(P.num_ranges + 255u) / 256u, 256, 0, ctx->stream>>>(P);
                KX_LAUNCHED(ctx);
                kxparse5::resolve_chunks_kernel<<<4 * ctx->sm_count, kxparse5::RES_WARPS * 32, 0, ctx->stream>>>(P);
            } else if (version == 4) {
                kxparse4::Params4 P;
                P.text = d_text; P.n = n; P.base = base; P.num_chunks = num_tiles;
                P.num_sc = (num_tiles + kxparse4::SCC4 - 1) / kxparse4::SCC4;
                P.num_ranges = (P.num_sc + kxparse4::RSC - 1) / kxparse4::RSC;
                P.range_state = t->tile_state;                               // [num_ranges]
                P.deferred = (uint32_t *)(t->tile_state + P.num_ranges);     // [num_tiles]
                P.tab = t->dev; P.carry_in = carry_in;
                P.tma_limit = n >= (size_t)kxparse2::STG_BYTES ? (uint32_t)((n - kxparse2::STG_BYTES) / kxparse2::CW) + 1u : 0u;
                kxparse4::parse_kernel_v4<<<parse_grid_v4(ctx, P.num_ranges), kxparse2::NT, sizeof(kxparse4::CtaSmem4), ctx->stream>>>(P);
                KX_LAUNCHED(ctx);
                kxparse4::resolve_deferred_kernel<<<2 * ctx->sm_count, kxparse4::RES_WARPS * 32, 0, ctx->stream>>>(P);
            } else if (version == 3) {
                kxparse3::Params3 P;
                P.text = d_text; P.n = n; P.base = base; P.num_chunks = num_tiles; P.num_sc = num_sc;
                P.sc_state = t->tile_state; P.tab = t->dev; P.carry_in = carry_in;
                kxparse3::parse_kernel_v3<<<parse_grid_v3(ctx, num_sc), kxparse2::NT, sizeof(kxparse3::CtaSmem3), ctx->stream>>>(P);
            } else if (version == 2) {
                kxparse2::Params P;
                P.text = d_text; P.n = n; P.base = base; P.num_chunks = num_tiles;
                P.chunk_state = t->tile_state; P.tab = t->dev; P.carry_in = carry_in;
                const size_t smem = sizeof(kxparse2::WarpSmem) * kxparse2::WARPS;
                kxparse2::parse_kernel_v2<<<parse_grid_v2(ctx, num_tiles), kxparse2::NT, smem, ctx->stream>>>(P);
            } else {
                kxparse::ParseParams P;
                P.text = d_text; P.n = n; P.base = base; P.num_tiles = num_tiles;
                P.tile_state = t->tile_state; P.tab = t->dev; P.carry_in = carry_in;
                kxparse::parse_kernel<<<parse_grid(ctx, num_tiles), kxparse::NT, sizeof(kxparse::ParseSmem), ctx->stream>>>(P);
            }
            KX_LAUNCHED(ctx);
        }
        {
            KxTimer tm(ctx, KXPU_T_FINALIZE);
            rc = launch_finalize(ctx, t, d_text, n, base, check_valid);
        }
        if (rc != KXPU_OK) { table_release(ctx, t); return rc; }
        // the join does not need anything from the host: enqueue it before the round trip below
        // (it is simply run again if the table has to be rebuilt)
        if (join) launch_lookup(ctx, t, join->d_keys, join->n, join->d_rows);
        cudaMemcpyAsync(ctx->h_ctl, t->dev.counters, KX_C_COUNT * 4, cudaMemcpyDeviceToHost, ctx->stream);
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            KX_SET_ERR(ctx, "parse/finalize failed: %s", cudaGetErrorString(e));
            table_release(ctx, t);
            return KXPU_E_CUDA;
        }
        if (!check_valid && ctx->h_ctl[KX_C_LONGLINE_HINT]) {
            // sharded load: the shard's bufio.ErrTooLong cut-off travels in the slab header
            kxparse::trunc_kernel<<<1, 1024, 0, ctx->stream>>>(d_text, n, base, t->dev.trunc);
            KX_LAUNCHED(ctx);
            KX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        }
        if (ctx->h_ctl[KX_C_NEED_TRUNC] == 2u) {
            // a >= 1 KiB stretch without a line start was seen: compute the exact
            // bufio.ErrTooLong cut-off and finalize again.
            kxparse::trunc_kernel<<<1, 1024, 0, ctx->stream>>>(d_text, n, base, t->dev.trunc);
            KX_LAUNCHED(ctx);
            uint32_t one = 1;
            cudaMemcpyAsync(&t->dev.counters[KX_C_NEED_TRUNC], &one, 4, cudaMemcpyHostToDevice, ctx->stream);
            cudaMemsetAsync(&t->dev.counters[KX_C_NROWS], 0, 12, ctx->stream);  // NROWS, BLOB_CURSOR, BLOB_OVERFLOW
            cudaMemsetAsync(&t->dev.counters[KX_C_NSEL], 0, 4, ctx->stream);
            rc = launch_finalize(ctx, t, d_text, n, base, check_valid);
            if (rc != KXPU_OK) { table_release(ctx, t); return rc; }
            if (join) launch_lookup(ctx, t, join->d_keys, join->n, join->d_rows);
            cudaMemcpyAsync(ctx->h_ctl, t->dev.counters, KX_C_COUNT * 4, cudaMemcpyDeviceToHost, ctx->stream);
            KX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        }
        if (ctx->h_ctl[KX_C_PEND_OVERFLOW]) {
            // more than PENDCAP device lines of one 64 KiB super-chunk are governed by an earlier
            // super-chunk (lines shorter than ~30 bytes): the warp-autonomous kernel has no such limit
            table_release(ctx, t);
            version = 2;
            continue;
        }
        if (ctx->h_ctl[KX_C_OVERFLOW] || ctx->h_ctl[KX_C_NKEYS] > t->dev.max_keys) {
            table_release(ctx, t);
            if (cap >= (1u << 28)) return KXPU_E_CAPACITY;
            cap <<= 2;
            continue;
        }
        if (ctx->h_ctl[KX_C_BLOB_OVERFLOW]) {
            table_release(ctx, t);
            if ((size_t)blob_cap >= n) return KXPU_E_CAPACITY;
            blob_cap = (uint32_t)std::min<size_t>(n, (size_t)blob_cap * 8);
            continue;
        }
        t->n_rows = ctx->h_ctl[KX_C_NSEL];  // row handle = index of the slot in the selection
        t->blob_used = ctx->h_ctl[KX_C_BLOB_CURSOR];
        *out = t;
        return KXPU_OK;
    }
    return KXPU_E_CAPACITY;
}

// ------------------------------------------------------------------ sharded load: merge side
namespace kxmerge {
using namespace kxcomm;

// every gathered candidate row / vendor row folds into the merged table with the parse rule
__global__ void __launch_bounds__(256) merge_insert_kernel(const uint8_t *gather, int R, size_t stride, SlabCaps caps,
                                                           KxTableDev tab, const uint32_t *peer_timeout) {
    const uint32_t per = caps.rows > caps.vendors ? caps.rows : caps.vendors;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int r = (int)(t / per);
    const uint32_t i = (uint32_t)(t % per);
    if (r >= R) return;
    const uint8_t *slab = gather + (size_t)r * stride;
    const SlabHeader *h = reinterpret_cast<const SlabHeader *>(slab);
    if (i == 0 && h->trunc != KX_NO_OFF) atomicMin(tab.trunc, h->trunc);
    if (i == 0 && h->overflow) atomicOr(&tab.counters[KX_C_SLAB_OVERFLOW], 1u);
    if (t == 0 && peer_timeout && *peer_timeout) atomicOr(&tab.counters[KX_C_SLAB_OVERFLOW], 2u);  // a peer never delivered  // some rank outgrew the slab: the host retries with larger ones
    if (i < h->n_rows) {
        const SlabRow row = reinterpret_cast<const SlabRow *>(slab + slab_rows_off())[i];
        kxparse2::table_fold(tab, row.key, row.line, row.anchor);
    }
    if (i < h->n_vendors) {
        const SlabVendor v = reinterpret_cast<const SlabVendor *>(slab + slab_vendors_off(caps))[i];
        atomicMin(&tab.vendor_first[v.vendor & 0xffffu], v.first);
    }
}

__global__ void __launch_bounds__(256) merge_finalize_kernel(KxTableDev tab, int32_t *row_of_slot, uint32_t *row_key,
                                                             unsigned long long *row_line, unsigned long long *row_anchor,
                                                             uint32_t *row_name_off, uint32_t *row_name_len) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot > tab.cap) return;
    const unsigned long long line = tab.min_line[slot];
    const uint32_t key = slot == tab.cap ? KX_EMPTY_KEY : tab.keys[slot];
    int32_t row = -1;
    if (line != KX_NO_OFF && !(slot < tab.cap && key == KX_EMPTY_KEY)) {
        const unsigned long long anchor = tab.min_anchor[slot];
        if (anchor == tab.vendor_first[key >> 16] && line < *tab.trunc) {
            row = (int32_t)atomicAdd(&tab.counters[KX_C_NROWS], 1u);
            row_key[row] = key; row_line[row] = line; row_anchor[row] = anchor;
            row_name_off[row] = 0; row_name_len[row] = 0;
        }
    }
    row_of_slot[slot] = row;
}

// the winning row's name stays where the all-gather put it: record its offset in the gather buffer
__global__ void __launch_bounds__(256) merge_names_kernel(const uint8_t *gather, int R, size_t stride, SlabCaps caps,
                                                          KxTableDev tab, const int32_t *row_of_slot,
                                                          uint32_t *row_name_off, uint32_t *row_name_len, uint8_t *own_blob,
                                                          uint32_t own_cap) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int r = (int)(t / caps.rows);
    const uint32_t i = (uint32_t)(t % caps.rows);
    if (r >= R) return;
    const uint8_t *slab = gather + (size_t)r * stride;
    const SlabHeader *h = reinterpret_cast<const SlabHeader *>(slab);
    if (i >= h->n_rows) return;
    const SlabRow row = reinterpret_cast<const SlabRow *>(slab + slab_rows_off())[i];
    uint32_t slot;
    if (row.key == KX_EMPTY_KEY) slot = tab.cap;
    else {
        slot = kx_hash(row.key) >> tab.shift;
        for (uint32_t step = 0; step < tab.cap; step++) {
            const uint32_t k = tab.keys[slot];
            if (k == row.key) break;
            if (k == KX_EMPTY_KEY) return;
            slot = (slot + 1) & (tab.cap - 1);
        }
    }
    const int32_t out = row_of_slot[slot];
    if (out >= 0 && tab.min_line[slot] == row.line) {
        const size_t src = (size_t)r * stride + slab_blob_off(caps) + row.name_off;
        if (own_blob) {
            // the gather buffer is reused by the next load (peer-memory path): the table keeps its own copy
            const uint32_t at = row.name_len ? atomicAdd(&tab.counters[KX_C_BLOB_CURSOR], row.name_len) : 0u;
            if (at + row.name_len > own_cap) { tab.counters[KX_C_BLOB_OVERFLOW] = 1u; return; }
            for (uint32_t b = 0; b < row.name_len; b++) own_blob[at + b] = gather[src + b];
            row_name_off[out] = at;
        } else {
            row_name_off[out] = (uint32_t)src;
        }
        row_name_len[out] = row.name_len;
    }
}
}  // namespace kxmerge

// own_names: copy the winning names into the table's own blob and leave d_gather to the caller
// (peer-memory path); otherwise the table takes d_gather over and serves the names from it.
int32_t kx_table_from_gather(kxpu_ctx *ctx, void *d_gather, int R, size_t stride, kxcomm::SlabCaps caps, kxpu_table **out,
                             bool own_names, const uint32_t *peer_timeout) {
    using namespace kxcomm;
    auto drop_gather = [&]() { if (!own_names) cudaFreeAsync(d_gather, ctx->stream); };
    // No host round trip for the slab headers: the insert kernel flags a slab overflow of any rank
    // in the counters that are read back anyway (every rank sees the same headers and retries alike).
    cudaError_t e = cudaSuccess;
    if ((size_t)R * stride >= 0xFFFFFFFFull) { drop_gather(); return KXPU_E_UNSUPPORTED; }
    uint32_t cap = 1u << 16;
    for (int attempt = 0; attempt < 8; attempt++) {
        kxpu_table *t = nullptr;
        const uint32_t own_cap = own_names ? (uint32_t)std::min<size_t>((size_t)R * caps.blob, 0xF0000000u) : 16u;
        int32_t rc = table_alloc(ctx, cap, own_cap, 1, &t);
        if (rc != KXPU_OK) { drop_gather(); return rc; }
        const uint32_t per = caps.rows > caps.vendors ? caps.rows : caps.vendors;
        const size_t nthreads = (size_t)R * per;
        kxmerge::merge_insert_kernel<<<(unsigned)((nthreads + 255) / 256), 256, 0, ctx->stream>>>((const uint8_t *)d_gather, R, stride,
                                                                                                caps, t->dev, peer_timeout);
        kxmerge::merge_finalize_kernel<<<(cap + 1 + 255) / 256, 256, 0, ctx->stream>>>(t->dev, t->row_of_slot, t->row_key, t->row_line,
                                                                                     t->row_anchor, t->row_name_off, t->row_name_len);
        kxmerge::merge_names_kernel<<<(unsigned)(((size_t)R * caps.rows + 255) / 256), 256, 0, ctx->stream>>>(
            (const uint8_t *)d_gather, R, stride, caps, t->dev, t->row_of_slot, t->row_name_off, t->row_name_len,
            own_names ? t->blob : nullptr, own_cap);
        ctx->launches += 3;
        cudaMemcpyAsync(ctx->h_ctl, t->dev.counters, KX_C_COUNT * 4, cudaMemcpyDeviceToHost, ctx->stream);
        e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            KX_SET_ERR(ctx, "merge failed: %s", cudaGetErrorString(e));
            table_release(ctx, t);
            drop_gather();
            return KXPU_E_CUDA;
        }
        if (ctx->h_ctl[KX_C_SLAB_OVERFLOW]) {
            const bool dead_peer = (ctx->h_ctl[KX_C_SLAB_OVERFLOW] & 2u) != 0;
            table_release(ctx, t);
            drop_gather();
            if (dead_peer) { KX_SET_ERR(ctx, "peer-memory exchange: a rank did not deliver its slab"); return KXPU_E_NCCL; }
            return KXPU_E_CAPACITY;
        }
        if (ctx->h_ctl[KX_C_OVERFLOW] || ctx->h_ctl[KX_C_NKEYS] > t->dev.max_keys) {
            table_release(ctx, t);
            if (cap >= (1u << 28)) { drop_gather(); return KXPU_E_UNSUPPORTED; }
            cap <<= 2;
            continue;
        }
        t->n_rows = ctx->h_ctl[KX_C_NROWS];
        if (own_names) {
            if (ctx->h_ctl[KX_C_BLOB_OVERFLOW]) { table_release(ctx, t); return KXPU_E_UNSUPPORTED; }
            t->blob_used = ctx->h_ctl[KX_C_BLOB_CURSOR];
        } else {
            t->gather = d_gather;               // names are served from the gathered slabs
            t->blob = (uint8_t *)d_gather;
            t->blob_cap = (uint32_t)((size_t)R * stride);
        }
        *out = t;
        return KXPU_OK;
    }
    drop_gather();
    return KXPU_E_UNSUPPORTED;
}

void kx_table_local_view(kxpu_table *t, KxTableDev *dev, uint32_t *cap, uint32_t *n_rows, uint32_t *blob_used,
                         const uint32_t **row_key, const unsigned long long **row_line, const unsigned long long **row_anchor,
                         const uint32_t **row_name_off, const uint32_t **row_name_len, const uint8_t **blob) {
    *dev = t->dev; *cap = t->cap; *n_rows = t->n_rows; *blob_used = t->blob_used;
    *row_key = t->row_key; *row_line = t->row_line; *row_anchor = t->row_anchor;
    *row_name_off = t->row_name_off; *row_name_len = t->row_name_len; *blob = t->blob;
}

void kx_table_release(kxpu_ctx *ctx, kxpu_table *t) { table_release(ctx, t); }

extern "C" int32_t kxpu_pciids_load_device(kxpu_ctx *ctx, const void *d_text, size_t n, kxpu_table **out) {
    KX_ENTER(ctx);
    if (!out || (!d_text && n)) return KXPU_E_INVALID;
    kx_clear_timings(ctx);
    return kx_build_table(ctx, (const uint8_t *)d_text, n, 0, 0, 1, out);
}

extern "C" int32_t kxpu_pciids_join_device(kxpu_ctx *ctx, const void *d_text, size_t n, const uint32_t *d_keys, size_t nq,
                                           int32_t *d_rows_out, kxpu_table **out) {
    KX_ENTER(ctx);
    if (!out || (!d_text && n) || (nq && (!d_keys || !d_rows_out))) return KXPU_E_INVALID;
    kx_clear_timings(ctx);
    const KxJoin j{d_keys, nq, d_rows_out};
    return kx_build_table_join(ctx, (const uint8_t *)d_text, n, 0, 0, 1, out, &j);
}

extern "C" int32_t kxpu_pciids_load(kxpu_ctx *ctx, const uint8_t *text, size_t n, kxpu_table **out) {
    KX_ENTER(ctx);
    if (!out || (!text && n)) return KXPU_E_INVALID;
    kx_clear_timings(ctx);
    void *d = nullptr;
    KX_CUDA(ctx, cudaMallocAsync(&d, n ? n : 16, ctx->stream));
    cudaError_t e = cudaMemcpyAsync(d, text, n, cudaMemcpyHostToDevice, ctx->stream);
    int32_t rc = e == cudaSuccess ? kx_build_table(ctx, (const uint8_t *)d, n, 0, 0, 1, out) : KXPU_E_CUDA;
    if (e != cudaSuccess) KX_SET_ERR(ctx, "H2D copy of the text failed: %s", cudaGetErrorString(e));
    cudaFreeAsync(d, ctx->stream);
    return rc;
}

extern "C" int32_t kxpu_table_free(kxpu_ctx *ctx, kxpu_table *t) {
    KX_ENTER(ctx);
    table_release(ctx, t);
    return KXPU_OK;
}

extern "C" int32_t kxpu_table_rows(kxpu_ctx *ctx, kxpu_table *t, uint32_t *n_rows) {
    if (!ctx || !t || !n_rows) return KXPU_E_INVALID;
    *n_rows = t->n_rows;
    return KXPU_OK;
}

extern "C" int32_t kxpu_table_export(kxpu_ctx *ctx, kxpu_table *t, uint32_t *keys, uint64_t *line_off, int32_t *rows,
                                     size_t cap, uint32_t *n_rows) {
    KX_ENTER(ctx);
    if (!t || !n_rows) return KXPU_E_INVALID;
    *n_rows = t->n_rows;
    if (cap < t->n_rows) return KXPU_E_NOSPACE;
    if (t->n_rows == 0) return KXPU_OK;
    if (!keys || !line_off || !rows) return KXPU_E_INVALID;
    std::vector<uint32_t> k(t->n_rows);
    std::vector<unsigned long long> l(t->n_rows);
    KX_CUDA(ctx, cudaMemcpyAsync(k.data(), t->row_key, t->n_rows * 4ull, cudaMemcpyDeviceToHost, ctx->stream));
    KX_CUDA(ctx, cudaMemcpyAsync(l.data(), t->row_line, t->n_rows * 8ull, cudaMemcpyDeviceToHost, ctx->stream));
    KX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    std::vector<int32_t> order(t->n_rows);
    for (uint32_t i = 0; i < t->n_rows; i++) order[i] = (int32_t)i;
    std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return l[a] < l[b]; });  // file order
    for (uint32_t i = 0; i < t->n_rows; i++) {
        keys[i] = k[order[i]];
        line_off[i] = l[order[i]];
        rows[i] = order[i];
    }
    return KXPU_OK;
}

// ------------------------------------------------------------------ lookup / names
static int32_t launch_lookup(kxpu_ctx *ctx, kxpu_table *t, const uint32_t *d_keys, size_t n, int32_t *d_rows) {
    if (n == 0) return KXPU_OK;
    KxTimer tm(ctx, KXPU_T_LOOKUP);
    size_t blocks = (n + 255) / 256;
    size_t maxb = (size_t)ctx->sm_count * 32;
    if (blocks > maxb) blocks = maxb;
    kxparse::lookup_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(d_keys, n, t->dev.keys, t->row_of_slot, t->cap,
                                                                     t->shift, d_rows);
    KX_LAUNCHED(ctx);
    KX_CUDA(ctx, cudaGetLastError());
    return KXPU_OK;
}

extern "C" int32_t kxpu_lookup_device(kxpu_ctx *ctx, kxpu_table *t, const uint32_t *d_keys, size_t n,
                                      int32_t *d_rows_out) {
    KX_ENTER(ctx);
    if (!t || (n && (!d_keys || !d_rows_out))) return KXPU_E_INVALID;
    kx_clear_timings(ctx);
    return launch_lookup(ctx, t, d_keys, n, d_rows_out);
}

extern "C" int32_t kxpu_lookup(kxpu_ctx *ctx, kxpu_table *t, const uint32_t *keys, size_t n, int32_t *rows_out) {
    KX_ENTER(ctx);
    if (!t || (n && (!keys || !rows_out))) return KXPU_E_INVALID;
    kx_clear_timings(ctx);
    if (n == 0) return KXPU_OK;
    uint32_t *d_keys = nullptr;
    int32_t *d_rows = nullptr;
    KX_CUDA(ctx, cudaMallocAsync((void **)&d_keys, n * 4, ctx->stream));
    KX_CUDA(ctx, cudaMallocAsync((void **)&d_rows, n * 4, ctx->stream));
    cudaMemcpyAsync(d_keys, keys, n * 4, cudaMemcpyHostToDevice, ctx->stream);
    int32_t rc = launch_lookup(ctx, t, d_keys, n, d_rows);
    cudaMemcpyAsync(rows_out, d_rows, n * 4, cudaMemcpyDeviceToHost, ctx->stream);
    cudaFreeAsync(d_keys, ctx->stream);
    cudaFreeAsync(d_rows, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (rc == KXPU_OK && e != cudaSuccess) { KX_SET_ERR(ctx, "lookup failed: %s", cudaGetErrorString(e)); rc = KXPU_E_CUDA; }
    return rc;
}

extern "C" int32_t kxpu_names(kxpu_ctx *ctx, kxpu_table *t, const int32_t *rows, size_t n, uint8_t *out, size_t cap,
                              uint32_t *offsets, size_t *need) {
    KX_ENTER(ctx);
    if (!t || !offsets || (n && !rows)) return KXPU_E_INVALID;
    kx_clear_timings(ctx);
    if (n == 0) { offsets[0] = 0; if (need) *need = 0; return KXPU_OK; }
    int32_t *d_rows = nullptr;
    uint32_t *d_lens = nullptr, *d_offs = nullptr;
    unsigned long long *d_part = nullptr;
    uint8_t *d_out = nullptr;
    const size_t np = kxscan::scratch_items(n + 1);
    KX_CUDA(ctx, cudaMallocAsync((void **)&d_rows, n * 4, ctx->stream));
    KX_CUDA(ctx, cudaMallocAsync((void **)&d_lens, (n + 1) * 4, ctx->stream));
    KX_CUDA(ctx, cudaMallocAsync((void **)&d_offs, (n + 1) * 4, ctx->stream));
    KX_CUDA(ctx, cudaMallocAsync((void **)&d_part, (np + 1) * 8, ctx->stream));
    cudaMemcpyAsync(d_rows, rows, n * 4, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemsetAsync(d_lens + n, 0, 4, ctx->stream);
    int32_t rc = KXPU_OK;
    {
        KxTimer tm(ctx, KXPU_T_NAMES);
        kxparse::name_len_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_rows, n, t->row_name_len,
                                                                                       t->n_rows, d_lens);
        KX_LAUNCHED(ctx);
        // scanning n+1 items makes offsets[n] the total
        kxscan::exclusive_scan<uint32_t>(ctx, d_lens, n + 1, d_offs, d_part, d_part + np);
    }
    cudaMemcpyAsync(offsets, d_offs, (n + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { KX_SET_ERR(ctx, "names failed: %s", cudaGetErrorString(e)); rc = KXPU_E_CUDA; }
    size_t total = rc == KXPU_OK ? offsets[n] : 0;
    if (need) *need = total;
    if (rc == KXPU_OK && total > cap) rc = KXPU_E_NOSPACE;
    if (rc == KXPU_OK && total > 0) {
        if (!out) rc = KXPU_E_INVALID;
        else {
            e = cudaMallocAsync((void **)&d_out, total, ctx->stream);
            if (e == cudaSuccess) {
                kxparse::name_copy_kernel<<<(unsigned)((n * 8 + 255) / 256), 256, 0, ctx->stream>>>(
                    d_rows, n, t->row_name_off, t->row_name_len, t->n_rows, t->blob, d_offs, d_out, total);
                KX_LAUNCHED(ctx);
                cudaMemcpyAsync(out, d_out, total, cudaMemcpyDeviceToHost, ctx->stream);
                cudaFreeAsync(d_out, ctx->stream);
                e = cudaStreamSynchronize(ctx->stream);
            }
            if (e != cudaSuccess) { KX_SET_ERR(ctx, "names copy failed: %s", cudaGetErrorString(e)); rc = KXPU_E_CUDA; }
        }
    }
    cudaFreeAsync(d_rows, ctx->stream);
    cudaFreeAsync(d_lens, ctx->stream);
    cudaFreeAsync(d_offs, ctx->stream);
    cudaFreeAsync(d_part, ctx->stream);
    return rc;
}
