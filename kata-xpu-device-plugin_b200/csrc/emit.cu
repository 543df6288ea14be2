// emit.cu -- K6 CDI spec emit (YAML / JSON), K7 Allocate names, ListAndWatch wire bytes.
//
// Reference: generateCDISpec (pkg/device_plugin/device_plugin.go:55-80), CdiSpec.Save
// (cdi/spec.go:85-127), updateResponseForCDI / QualifiedName
// (pkg/device_plugin/generic_device_plugin.go:274-299, cdi/cdi-utils.go:9) and the
// ListAndWatchResponse send (generic_device_plugin.go:224).
//
// Every output is a concatenation of per-device fragments whose length depends on the data
// (decimal widths, the YAML quoting predicate).  The CDI emitter is ONE kernel: a CTA takes a tile
// of 128 devices, computes the fragment lengths, scans them, gets the tile's output offset by a
// decoupled look-back over the tile aggregates (scan.cuh), builds the tile's bytes in shared memory
// (one warp per device, literal segments copied from a shared-memory pool) at the same 16-byte
// phase as their destination, and writes the tile with ONE TMA bulk store (cp.async.bulk
// shared -> global) plus at most 15 head / tail bytes.  Document header and tail belong to the first
// and the last tile.  Allocate names and ListAndWatch bytes (<= 24 B per item) are length -> single-
// pass scan -> thread-per-item write.
#include <vector>

#include "common.cuh"
#include "scan.cuh"

namespace kxemit {

constexpr int TILE = 128;      // devices per CTA
constexpr int EMIT_THREADS = 256;
constexpr int MAX_FRAG = 368;  // upper bound of one fragment: literals (<= 284) + 2 x 20 + 2 x 10 + 15 + 2 + slack; with the rest of
                               // TileSmem this keeps a CTA under 56.7 KB: four CTAs per SM, the 512 tiles of cfg5 are ONE wave
constexpr int POOL_MAX = 640;

// ------------------------------------------------------------------ templates
// YAML (yaml.v3, indent 2) and JSON (MarshalIndent "  ") literals between the variable
// fields: see SURVEY.md 8a-fmt for the derivation.
#define KX_Y0 "  - name: \""
#define KX_Y1 "\"\n    annotations:\n      attach-pci: \"true\"\n      bdf: "
#define KX_Y2 "\n      cdi.k8s.io/vfio"
#define KX_Y3 ": nvidia.com/gpu="
#define KX_Y4 "\n    containerEdits:\n      deviceNodes:\n        - path: /dev/vfio/"
#define KX_Y5 "\n"
#define KX_YH "cdiVersion: 0.6.0\nkind: nvidia.com/gpu\ndevices:\n"
#define KX_YT ""
#define KX_J0 "    {\n      \"name\": \""
#define KX_J1 "\",\n      \"annotations\": {\n        \"attach-pci\": \"true\",\n        \"bdf\": \""
#define KX_J2 "\",\n        \"cdi.k8s.io/vfio"
#define KX_J3 "\": \"nvidia.com/gpu="
#define KX_J4 "\"\n      },\n      \"containerEdits\": {\n        \"deviceNodes\": [\n          {\n            \"path\": \"/dev/vfio/"
#define KX_J5 "\"\n          }\n        ]\n      }\n    }"
#define KX_JH "{\n  \"cdiVersion\": \"0.6.0\",\n  \"kind\": \"nvidia.com/gpu\",\n  \"devices\": [\n"
#define KX_JT "  ],\n  \"containerEdits\": {}\n}"
// pool = six literals | document head | document tail
__constant__ char c_yaml_pool[] = KX_Y0 KX_Y1 KX_Y2 KX_Y3 KX_Y4 KX_Y5 KX_YH KX_YT;
__constant__ char c_json_pool[] = KX_J0 KX_J1 KX_J2 KX_J3 KX_J4 KX_J5 KX_JH KX_JT;
static const char *h_yaml_parts[8] = {KX_Y0, KX_Y1, KX_Y2, KX_Y3, KX_Y4, KX_Y5, KX_YH, KX_YT};
static const char *h_json_parts[8] = {KX_J0, KX_J1, KX_J2, KX_J3, KX_J4, KX_J5, KX_JH, KX_JT};

static const char h_yaml_empty[] = "cdiVersion: 0.6.0\nkind: nvidia.com/gpu\ndevices: []\n";
static const char h_json_empty[] =
    "{\n  \"cdiVersion\": \"0.6.0\",\n  \"kind\": \"nvidia.com/gpu\",\n  \"devices\": null,\n  \"containerEdits\": {}\n}";

__device__ __forceinline__ uint32_t dec_len(unsigned long long v) {
    uint32_t l = 1;
    while (v >= 10ull) { v /= 10ull; l++; }
    return l;
}
__device__ __forceinline__ void dec_write(unsigned long long v, uint32_t len, uint8_t *dst) {
    for (uint32_t k = len; k > 0; k--) { dst[k - 1] = (uint8_t)('0' + (uint32_t)(v % 10ull)); v /= 10ull; }
}
__device__ __forceinline__ uint32_t bdf_len16(const uint8_t *b) {
    uint32_t l = 0;
    while (l < 16u && b[l]) l++;
    return l;
}
// yaml.v3 isBase60Float: ^[-+]?[0-9][0-9_]*(?::[0-5]?[0-9])+(?:\.[0-9_]*)?$  (resolve.go);
// such a string would be read back as a sexagesimal number, so encode.go quotes it.
__device__ __forceinline__ bool is_base60(const uint8_t *s, uint32_t len) {
    uint32_t i = 0;
    auto dig = [](uint8_t c) { return c >= '0' && c <= '9'; };
    if (i < len && (s[i] == '-' || s[i] == '+')) i++;
    if (!(i < len && dig(s[i]))) return false;
    i++;
    while (i < len && (dig(s[i]) || s[i] == '_')) i++;
    uint32_t groups = 0;
    while (i < len && s[i] == ':') {
        uint32_t j = i + 1;
        if (!(j < len && dig(s[j]))) break;
        if (j + 1 < len && dig(s[j + 1])) j += s[j] <= '5' ? 2u : 1u;
        else j += 1;
        i = j; groups++;
    }
    if (!groups) return false;
    if (i < len && s[i] == '.') { i++; while (i < len && (dig(s[i]) || s[i] == '_')) i++; }
    return i == len;
}
// bytes Go's encoders would escape or that change YAML plain-scalar rules are outside the
// supported domain; PCI addresses only use [0-9a-f:.]
__device__ __forceinline__ bool bdf_charset_ok(const uint8_t *s, uint32_t len) {
    if (len == 0) return false;
    for (uint32_t i = 0; i < len; i++) {
        uint8_t c = s[i];
        if (!((c >= '0' && c <= '9') || (c >= 'a' && c <= 'f') || c == ':' || c == '.')) return false;
    }
    return true;
}

struct EmitParams {
    const kxpu_cdidev *devs;
    uint32_t n;
    uint16_t off[8], len[8];  // literal k / head (6) / tail (7) inside the pool
    uint32_t pool_len, lit_total;
    uint8_t *out;
    unsigned long long *state;  // tile status words (scan.cuh look-back)
    uint32_t epoch;
    unsigned long long *total_out;
    uint32_t *flags;
};

struct TileSmem {
    alignas(16) uint8_t stage[TILE * MAX_FRAG + 512];
    uint8_t pool[POOL_MAX];
    uint8_t idx[TILE][20], grp[TILE][12], bdf[TILE][16];
    uint32_t meta[TILE];    // il | gl << 8 | bl << 16 | quoted << 24
    uint32_t start[TILE];   // fragment offset inside the tile
    unsigned long long base;
    uint32_t wsum[EMIT_THREADS / 32];
    uint32_t tile_total;
};

template <int FMT>
__global__ void __launch_bounds__(EMIT_THREADS) k_cdi_fused(const EmitParams E) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    TileSmem &S = *reinterpret_cast<TileSmem *>(smem_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
    const uint32_t tile = blockIdx.x, i0 = tile * TILE;
    const bool first_tile = tile == 0, last_tile = tile == gridDim.x - 1;
    const char *cpool = FMT == KXPU_FMT_YAML ? c_yaml_pool : c_json_pool;
    for (uint32_t k = tid; k < E.pool_len; k += EMIT_THREADS) S.pool[k] = (uint8_t)cpool[k];

    // ---- fragment lengths and variable fields: one thread per device
    uint32_t flen = 0;
    if (tid < TILE && i0 + tid < E.n) {
        const uint4 *p = reinterpret_cast<const uint4 *>(E.devs + i0 + tid);
        const uint4 q0 = p[0], q1 = p[1];
        const uint8_t *bdf = reinterpret_cast<const uint8_t *>(&q0);
        const uint32_t group = q1.x;
        const unsigned long long index = ((unsigned long long)q1.w << 32) | q1.z;
        const uint32_t bl = bdf_len16(bdf), il = dec_len(index), gl = dec_len(group);
        if (!bdf_charset_ok(bdf, bl)) E.flags[0] = 1u;
        const bool quoted = FMT == KXPU_FMT_YAML && is_base60(bdf, bl);
        dec_write(index, il, S.idx[tid]);
        dec_write(group, gl, S.grp[tid]);
        *reinterpret_cast<uint4 *>(S.bdf[tid]) = q0;
        S.meta[tid] = il | (gl << 8) | (bl << 16) | ((quoted ? 1u : 0u) << 24);
        flen = E.lit_total + 2u * il + 2u * gl + bl + (quoted ? 2u : 0u);
        if (FMT == KXPU_FMT_JSON) flen += (i0 + tid + 1u < E.n) ? 2u : 1u;  // ",\n" between devices, "\n" after the last
    }
    // ---- scan of the 128 lengths (threads >= TILE contribute 0)
    uint32_t incl = kxscan::warp_incl(flen);
    if (lane == 31) S.wsum[w] = incl;
    __syncthreads();
    if (w == 0) {
        const uint32_t x = lane < EMIT_THREADS / 32 ? S.wsum[lane] : 0u;
        const uint32_t xi = kxscan::warp_incl(x);
        if (lane < EMIT_THREADS / 32) S.wsum[lane] = xi - x;
        if (lane == EMIT_THREADS / 32 - 1) S.tile_total = xi;
    }
    __syncthreads();
    const uint32_t head_len = first_tile ? E.len[6] : 0u, tail_len = last_tile ? E.len[7] : 0u;
    const uint32_t tile_total = S.tile_total;
    if (tid < TILE) S.start[tid] = head_len + S.wsum[w] + incl - flen;
    // ---- the tile's offset in the document: decoupled look-back over the tile aggregates
    if (w == 0) {
        const unsigned long long agg = (unsigned long long)head_len + tile_total + tail_len;
        const unsigned long long excl = kxscan::lookback(E.state, tile, agg, E.epoch);
        if (lane == 0) {
            S.base = excl;
            if (last_tile) *E.total_out = excl + agg;
        }
    }
    __syncthreads();
    const unsigned long long base = S.base;
    const uint32_t mis = (uint32_t)((reinterpret_cast<uintptr_t>(E.out) + base) & 15u);  // same 16-byte phase in smem and global
    uint8_t *stg = S.stage + mis;
    if (first_tile) for (uint32_t k = tid; k < head_len; k += EMIT_THREADS) stg[k] = S.pool[E.off[6] + k];
    if (last_tile) for (uint32_t k = tid; k < tail_len; k += EMIT_THREADS) stg[head_len + tile_total + k] = S.pool[E.off[7] + k];
    // ---- fragments: one warp per device, segment by segment
    for (uint32_t d = w; d < (uint32_t)TILE && i0 + d < E.n; d += EMIT_THREADS / 32) {
        const uint32_t m = S.meta[d];
        const uint32_t il = m & 0xffu, gl = (m >> 8) & 0xffu, bl = (m >> 16) & 0xffu;
        const bool quoted = (m >> 24) != 0u;
        uint8_t *dst = stg + S.start[d];
        uint32_t o = 0;
        auto lit = [&](int k) {
            const uint32_t L = E.len[k];
            const uint8_t *src = S.pool + E.off[k];
            for (uint32_t l = lane; l < L; l += 32u) dst[o + l] = src[l];
            o += L;
        };
        auto var = [&](const uint8_t *src, uint32_t L) {  // L <= 20
            if (lane < L) dst[o + lane] = src[lane];
            o += L;
        };
        auto quote = [&]() {
            if (FMT == KXPU_FMT_YAML && quoted) { if (lane == 0) dst[o] = (uint8_t)'"'; o += 1u; }
        };
        lit(0); var(S.idx[d], il); lit(1); quote(); var(S.bdf[d], bl); quote(); lit(2); var(S.grp[d], gl);
        lit(3); var(S.idx[d], il); lit(4); var(S.grp[d], gl); lit(5);
        if (FMT == KXPU_FMT_JSON) {
            const bool more = i0 + d + 1u < E.n;
            if (lane == 0) { if (more) { dst[o] = (uint8_t)','; dst[o + 1] = (uint8_t)'\n'; } else dst[o] = (uint8_t)'\n'; }
        }
    }
    // generic-proxy writes to shared memory must be visible to the async proxy (TMA) that reads them
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    // ---- one bulk store for the 16-byte aligned body, byte stores for the ragged ends
    const uint32_t total = head_len + tile_total + tail_len;
    uint8_t *g = E.out + base;
    const uint32_t lead = total < 16u ? total : ((16u - mis) & 15u);
    const uint32_t body = (total - lead) & ~15u;
    if (tid == 0 && body) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g + lead),
                     "r"((uint32_t)__cvta_generic_to_shared(stg + lead)), "r"(body)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if (tid >= 32 && tid < 32 + lead) g[tid - 32] = stg[tid - 32];
    const uint32_t rest = total - lead - body;
    if (tid >= 64 && tid < 64 + rest) g[lead + body + tid - 64] = stg[lead + body + tid - 64];
    if (tid == 0 && body) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // smem must outlive the read
}

// ------------------------------------------------------------------ Allocate names
__global__ void __launch_bounds__(256) k_alloc_len(const unsigned long long *__restrict__ idx, uint32_t n,
                                                   uint32_t *__restrict__ lens) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    lens[i] = i < n ? 15u + dec_len(idx[i]) : 0u;  // len("nvidia.com/gpu=") == 15
}
__global__ void __launch_bounds__(256) k_alloc_write(const unsigned long long *__restrict__ idx, uint32_t n,
                                                     const uint32_t *__restrict__ offs, uint8_t *__restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const char pre[16] = "nvidia.com/gpu=";
    uint8_t *dst = out + offs[i];
#pragma unroll
    for (int k = 0; k < 15; k++) dst[k] = (uint8_t)pre[k];
    unsigned long long v = idx[i];
    dec_write(v, dec_len(v), dst + 15);
}

// ------------------------------------------------------------------ ListAndWatchResponse
// repeated Device devices = 1; Device { string ID = 1; string health = 2; }
__global__ void __launch_bounds__(256) k_lw_len(const uint32_t *__restrict__ groups, const uint8_t *__restrict__ healthy,
                                                uint32_t n, uint32_t *__restrict__ lens) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    if (i == n) { lens[i] = 0; return; }
    uint32_t hl = (!healthy || healthy[i]) ? 7u : 9u;
    lens[i] = 2u + 2u + dec_len(groups[i]) + 2u + hl;
}
__global__ void __launch_bounds__(256) k_lw_write(const uint32_t *__restrict__ groups, const uint8_t *__restrict__ healthy,
                                                  uint32_t n, const uint32_t *__restrict__ offs, uint8_t *__restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool ok = !healthy || healthy[i];
    const char hs[10] = "Unhealthy";
    const uint32_t hl = ok ? 7u : 9u, gl = dec_len(groups[i]);
    uint8_t *d = out + offs[i];
    d[0] = 0x0a; d[1] = (uint8_t)(2u + gl + 2u + hl);
    d[2] = 0x0a; d[3] = (uint8_t)gl;
    dec_write(groups[i], gl, d + 4);
    d[4 + gl] = 0x12; d[5 + gl] = (uint8_t)hl;
    for (uint32_t k = 0; k < hl; k++) d[6 + gl + k] = (uint8_t)hs[ok ? k + 2 : k];  // "Healthy" = "Unhealthy"+2 with 'h'->'H'
    if (ok) d[6 + gl] = (uint8_t)'H';
}

}  // namespace kxemit

using namespace kxemit;

extern "C" int32_t kxpu_cdi_emit(kxpu_ctx *ctx, int32_t format, const kxpu_cdidev *devs, size_t n, uint8_t *out,
                                 size_t cap, size_t *len) {
    if (!ctx || !len || (n && !devs) || (format != KXPU_FMT_YAML && format != KXPU_FMT_JSON)) return KXPU_E_INVALID;
    if (n >= 0x7FFFFFFFull) return KXPU_E_UNSUPPORTED;
    std::lock_guard<std::mutex> guard(ctx->mu);
    cudaSetDevice(ctx->device);
    kx_clear_timings(ctx);
    if (n == 0) {  // Devices stays nil: yaml "devices: []", json "devices": null (cdi/spec.go:42-49)
        const char *doc = format == KXPU_FMT_YAML ? h_yaml_empty : h_json_empty;
        *len = strlen(doc);
        if (cap < *len || !out) return KXPU_E_NOSPACE;
        memcpy(out, doc, *len);
        return KXPU_OK;
    }
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(k_cdi_fused<KXPU_FMT_YAML>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileSmem));
        cudaFuncSetAttribute(k_cdi_fused<KXPU_FMT_JSON>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileSmem));
        attr_done = true;
    }
    const char **parts = format == KXPU_FMT_YAML ? h_yaml_parts : h_json_parts;
    EmitParams E;
    memset(&E, 0, sizeof E);
    uint32_t acc = 0;
    for (int k = 0; k < 8; k++) {
        E.off[k] = (uint16_t)acc;
        E.len[k] = (uint16_t)strlen(parts[k]);
        acc += E.len[k];
        if (k < 6) E.lit_total += E.len[k];
    }
    E.pool_len = acc;
    const uint32_t N = (uint32_t)n;
    const uint32_t tiles = (N + TILE - 1) / TILE;
    const size_t bound = (size_t)n * (E.lit_total + 2 * 20 + 2 * 10 + 15 + 2 + 2) + E.len[6] + E.len[7] + 64;  // no fragment is longer
    if (E.lit_total + 2 * 20 + 2 * 10 + 15 + 2 + 2 > (uint32_t)MAX_FRAG) return KXPU_E_INVALID;  // the literals grew: MAX_FRAG must follow
    KxScratch sc(ctx);
    kxpu_cdidev *d_devs = nullptr;
    uint8_t *d_out = nullptr;
    unsigned long long *d_total = nullptr;
    KX_CUDA(ctx, sc.alloc((void **)&d_devs, n * sizeof(kxpu_cdidev)));
    KX_CUDA(ctx, sc.alloc((void **)&d_out, bound));
    KX_CUDA(ctx, sc.alloc((void **)&d_total, 16));
    cudaMemcpyAsync(d_devs, devs, n * sizeof(kxpu_cdidev), cudaMemcpyHostToDevice, ctx->stream);
    cudaMemsetAsync(d_total, 0, 16, ctx->stream);
    E.devs = d_devs; E.n = N; E.out = d_out; E.total_out = d_total; E.flags = (uint32_t *)(d_total + 1);
    E.state = kx_scan_state(ctx, tiles);
    if (!E.state) return KXPU_E_NOMEM;
    E.epoch = kx_next_epoch(ctx);
    {
        KxTimer tm(ctx, KXPU_T_EMIT);
        if (format == KXPU_FMT_YAML) k_cdi_fused<KXPU_FMT_YAML><<<tiles, EMIT_THREADS, sizeof(TileSmem), ctx->stream>>>(E);
        else k_cdi_fused<KXPU_FMT_JSON><<<tiles, EMIT_THREADS, sizeof(TileSmem), ctx->stream>>>(E);
        KX_LAUNCHED(ctx);
    }
    unsigned long long h[2] = {0, 0};
    cudaMemcpyAsync(h, d_total, 16, cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { KX_SET_ERR(ctx, "cdi_emit failed: %s", cudaGetErrorString(e)); return KXPU_E_CUDA; }
    if ((uint32_t)h[1]) { KX_SET_ERR(ctx, "cdi_emit: bdf outside [0-9a-f:.]"); return KXPU_E_UNSUPPORTED; }
    const size_t total = (size_t)h[0];
    *len = total;
    if (cap < total || !out) return KXPU_E_NOSPACE;
    cudaMemcpyAsync(out, d_out, total, cudaMemcpyDeviceToHost, ctx->stream);
    e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { KX_SET_ERR(ctx, "cdi_emit D2H failed: %s", cudaGetErrorString(e)); return KXPU_E_CUDA; }
    return KXPU_OK;
}

// shared driver of the two "thread per item" emitters
template <typename LenK, typename WriteK>
static int32_t emit_items(kxpu_ctx *ctx, size_t n, size_t in_bytes, const void *h_in, const uint8_t *h_in2, uint8_t *out,
                          size_t cap, uint32_t *offsets, size_t *need, LenK lenk, WriteK writek) {
    const uint32_t N = (uint32_t)n;
    uint8_t *b = nullptr;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) / 256 * 256; return o; };
    size_t o_in = take(in_bytes), o_in2 = take(h_in2 ? n : 16), o_lens = take((n + 1) * 4), o_offs = take((n + 1) * 4);
    KxScratch sc(ctx);
    KX_CUDA(ctx, sc.alloc((void **)&b, off));
    cudaMemcpyAsync(b + o_in, h_in, in_bytes, cudaMemcpyHostToDevice, ctx->stream);
    if (h_in2) cudaMemcpyAsync(b + o_in2, h_in2, n, cudaMemcpyHostToDevice, ctx->stream);
    uint32_t *d_lens = (uint32_t *)(b + o_lens), *d_offs = (uint32_t *)(b + o_offs);
    lenk(b + o_in, h_in2 ? b + o_in2 : nullptr, N, d_lens);
    ctx->launches++;
    kxscan::exclusive_scan<uint32_t>(ctx, d_lens, n + 1, d_offs, nullptr);
    std::vector<uint32_t> tmp;
    uint32_t *h_offs = offsets;
    if (!h_offs) { tmp.resize(n + 1); h_offs = tmp.data(); }
    cudaMemcpyAsync(h_offs, d_offs, (n + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    int32_t rc = KXPU_OK;
    if (e != cudaSuccess) { KX_SET_ERR(ctx, "emit sizing failed: %s", cudaGetErrorString(e)); rc = KXPU_E_CUDA; }
    const size_t total = rc == KXPU_OK ? h_offs[n] : 0;
    if (need) *need = total;
    if (rc == KXPU_OK && (cap < total || (!out && total))) rc = KXPU_E_NOSPACE;
    if (rc == KXPU_OK && total) {
        uint8_t *d_out = nullptr;
        e = sc.alloc((void **)&d_out, total);
        if (e == cudaSuccess) {
            writek(b + o_in, h_in2 ? b + o_in2 : nullptr, N, d_offs, d_out);
            ctx->launches++;
            cudaMemcpyAsync(out, d_out, total, cudaMemcpyDeviceToHost, ctx->stream);
            e = cudaStreamSynchronize(ctx->stream);
        }
        if (e != cudaSuccess) { KX_SET_ERR(ctx, "emit write failed: %s", cudaGetErrorString(e)); rc = KXPU_E_CUDA; }
    }
    return rc;
}

extern "C" int32_t kxpu_alloc_names(kxpu_ctx *ctx, const uint64_t *idx, size_t n, uint8_t *out, size_t cap,
                                    uint32_t *offsets, size_t *need) {
    if (!ctx || !offsets || (n && !idx)) return KXPU_E_INVALID;
    if (n >= 0x7FFFFFFFull) return KXPU_E_UNSUPPORTED;
    std::lock_guard<std::mutex> guard(ctx->mu);
    cudaSetDevice(ctx->device);
    kx_clear_timings(ctx);
    if (n == 0) { offsets[0] = 0; if (need) *need = 0; return KXPU_OK; }
    cudaStream_t st = ctx->stream;
    return emit_items(
        ctx, n, n * 8, idx, nullptr, out, cap, offsets, need,
        [st](const uint8_t *in, const uint8_t *, uint32_t N, uint32_t *lens) {
            k_alloc_len<<<(N + 1 + 255) / 256, 256, 0, st>>>((const unsigned long long *)in, N, lens);
        },
        [st](const uint8_t *in, const uint8_t *, uint32_t N, const uint32_t *offs, uint8_t *o) {
            k_alloc_write<<<(N + 255) / 256, 256, 0, st>>>((const unsigned long long *)in, N, offs, o);
        });
}

extern "C" int32_t kxpu_lw_encode(kxpu_ctx *ctx, const uint32_t *group_ids, const uint8_t *healthy, size_t n,
                                  uint8_t *out, size_t cap, size_t *len) {
    if (!ctx || !len || (n && !group_ids)) return KXPU_E_INVALID;
    if (n >= 0x7FFFFFFFull) return KXPU_E_UNSUPPORTED;
    std::lock_guard<std::mutex> guard(ctx->mu);
    cudaSetDevice(ctx->device);
    kx_clear_timings(ctx);
    if (n == 0) { *len = 0; return KXPU_OK; }
    cudaStream_t st = ctx->stream;
    return emit_items(
        ctx, n, n * 4, group_ids, healthy, out, cap, nullptr, len,
        [st](const uint8_t *in, const uint8_t *in2, uint32_t N, uint32_t *lens) {
            k_lw_len<<<(N + 1 + 255) / 256, 256, 0, st>>>((const uint32_t *)in, in2, N, lens);
        },
        [st](const uint8_t *in, const uint8_t *in2, uint32_t N, const uint32_t *offs, uint8_t *o) {
            k_lw_write<<<(N + 255) / 256, 256, 0, st>>>((const uint32_t *)in, in2, N, offs, o);
        });
}
