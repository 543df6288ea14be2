// emit.cu -- K6 CDI spec emit (YAML / JSON), K7 Allocate names, ListAndWatch wire bytes.
//
// Reference: generateCDISpec (pkg/device_plugin/device_plugin.go:55-80), CdiSpec.Save
// (cdi/spec.go:85-127), updateResponseForCDI / QualifiedName
// (pkg/device_plugin/generic_device_plugin.go:274-299, cdi/cdi-utils.go:9) and the
// ListAndWatchResponse send (generic_device_plugin.go:224).
//
// Every output is a concatenation of per-device fragments whose length depends on the data
// (decimal widths, the YAML quoting predicate), so each emitter is: fragment length ->
// exclusive scan -> one warp per device writes its fragment.  A fragment is a fixed
// sequence of literal and variable segments; lanes resolve "which segment owns output
// byte k" from a per-warp prefix table in shared memory, so stores are contiguous.
#include <vector>

#include "common.cuh"
#include "scan.cuh"

namespace kxemit {

enum SegKind : uint8_t { LIT = 0, IDX = 1, GRP = 2, BDF = 3, QUOTE = 4, SEP = 5 };
struct Seg { uint8_t kind; uint16_t off, len; };  // LIT: [off,off+len) of the literal pool

constexpr int MAX_SEGS = 14;

// ------------------------------------------------------------------ templates
// YAML (yaml.v3, indent 2) and JSON (MarshalIndent "  ") literals between the variable
// fields: see SURVEY.md 8a-fmt for the derivation.
#define KX_Y0 "  - name: \""
#define KX_Y1 "\"\n    annotations:\n      attach-pci: \"true\"\n      bdf: "
#define KX_Y2 "\n      cdi.k8s.io/vfio"
#define KX_Y3 ": nvidia.com/gpu="
#define KX_Y4 "\n    containerEdits:\n      deviceNodes:\n        - path: /dev/vfio/"
#define KX_Y5 "\n"
#define KX_J0 "    {\n      \"name\": \""
#define KX_J1 "\",\n      \"annotations\": {\n        \"attach-pci\": \"true\",\n        \"bdf\": \""
#define KX_J2 "\",\n        \"cdi.k8s.io/vfio"
#define KX_J3 "\": \"nvidia.com/gpu="
#define KX_J4 "\"\n      },\n      \"containerEdits\": {\n        \"deviceNodes\": [\n          {\n            \"path\": \"/dev/vfio/"
#define KX_J5 "\"\n          }\n        ]\n      }\n    }"
__constant__ char c_yaml_pool[] = KX_Y0 KX_Y1 KX_Y2 KX_Y3 KX_Y4 KX_Y5;
__constant__ char c_json_pool[] = KX_J0 KX_J1 KX_J2 KX_J3 KX_J4 KX_J5;
static const char *h_yaml_lits[6] = {KX_Y0, KX_Y1, KX_Y2, KX_Y3, KX_Y4, KX_Y5};
static const char *h_json_lits[6] = {KX_J0, KX_J1, KX_J2, KX_J3, KX_J4, KX_J5};

static const char h_yaml_head[] = "cdiVersion: 0.6.0\nkind: nvidia.com/gpu\ndevices:\n";
static const char h_yaml_empty[] = "cdiVersion: 0.6.0\nkind: nvidia.com/gpu\ndevices: []\n";
static const char h_json_head[] = "{\n  \"cdiVersion\": \"0.6.0\",\n  \"kind\": \"nvidia.com/gpu\",\n  \"devices\": [\n";
static const char h_json_tail[] = "  ],\n  \"containerEdits\": {}\n}";
static const char h_json_empty[] =
    "{\n  \"cdiVersion\": \"0.6.0\",\n  \"kind\": \"nvidia.com/gpu\",\n  \"devices\": null,\n  \"containerEdits\": {}\n}";

struct Template {
    Seg segs[MAX_SEGS];
    int nsegs;
    uint32_t lit_total;  // sum of literal lengths
};

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
    int format;
    Template tpl;
    uint32_t *lens;       // [n+1]
    const unsigned long long *offs;  // [n+1] exclusive scan of lens
    uint8_t *out;
    unsigned long long head_len;
    uint32_t *flags;
};

__global__ void __launch_bounds__(256) k_cdi_len(const EmitParams E) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > E.n) return;
    if (i == E.n) { E.lens[i] = 0; return; }
    const uint4 *p = reinterpret_cast<const uint4 *>(E.devs + i);
    uint4 q0 = p[0], q1 = p[1];
    const uint8_t *bdf = reinterpret_cast<const uint8_t *>(&q0);
    const uint32_t group = q1.x;
    const unsigned long long index = ((unsigned long long)q1.w << 32) | q1.z;
    const uint32_t bl = bdf_len16(bdf);
    if (!bdf_charset_ok(bdf, bl)) E.flags[0] = 1u;
    uint32_t len = E.tpl.lit_total + 2u * dec_len(index) + 2u * dec_len(group) + bl;
    if (E.format == KXPU_FMT_YAML) len += is_base60(bdf, bl) ? 2u : 0u;
    else len += (i + 1u < E.n) ? 2u : 1u;  // ",\n" between devices, "\n" after the last
    E.lens[i] = len;
}

constexpr int EMIT_WARPS = 8;

__global__ void __launch_bounds__(EMIT_WARPS * 32) k_cdi_write(const EmitParams E) {
    __shared__ uint8_t s_idx[EMIT_WARPS][24], s_grp[EMIT_WARPS][12], s_bdf[EMIT_WARPS][16];
    __shared__ uint16_t s_start[EMIT_WARPS][MAX_SEGS + 1];
    const uint32_t lane = threadIdx.x & 31u, wl = threadIdx.x >> 5;
    const uint32_t i = blockIdx.x * EMIT_WARPS + wl;
    if (i >= E.n) return;
    const char *pool = E.format == KXPU_FMT_YAML ? c_yaml_pool : c_json_pool;
    if (lane == 0) {
        const kxpu_cdidev *d = E.devs + i;
        const uint8_t *bdf = reinterpret_cast<const uint8_t *>(d->bdf);
        const uint32_t bl = bdf_len16(bdf);
        const uint32_t il = dec_len(d->index), gl = dec_len(d->iommu_group);
        dec_write(d->index, il, s_idx[wl]);
        dec_write(d->iommu_group, gl, s_grp[wl]);
        for (uint32_t k = 0; k < 16u; k++) s_bdf[wl][k] = bdf[k];
        const bool quoted = E.format == KXPU_FMT_YAML && is_base60(bdf, bl);
        const uint32_t seplen = (i + 1u < E.n) ? 2u : 1u;
        uint32_t acc = 0;
        for (int s = 0; s < E.tpl.nsegs; s++) {
            s_start[wl][s] = (uint16_t)acc;
            const Seg sg = E.tpl.segs[s];
            uint32_t l = sg.kind == LIT ? sg.len : sg.kind == IDX ? il : sg.kind == GRP ? gl : sg.kind == BDF ? bl
                         : sg.kind == QUOTE ? (quoted ? 1u : 0u) : seplen;
            acc += l;
        }
        s_start[wl][E.tpl.nsegs] = (uint16_t)acc;
    }
    __syncwarp();
    const uint32_t total = s_start[wl][E.tpl.nsegs];
    uint8_t *dst = E.out + E.head_len + E.offs[i];
    for (uint32_t k = lane; k < total; k += 32u) {
        int s = 0;
#pragma unroll
        for (int t = 1; t < MAX_SEGS; t++)
            if (t < E.tpl.nsegs && k >= s_start[wl][t]) s = t;
        const Seg sg = E.tpl.segs[s];
        const uint32_t r = k - s_start[wl][s];
        uint8_t c;
        switch (sg.kind) {
            case LIT: c = (uint8_t)pool[sg.off + r]; break;
            case IDX: c = s_idx[wl][r]; break;
            case GRP: c = s_grp[wl][r]; break;
            case BDF: c = s_bdf[wl][r]; break;
            case QUOTE: c = (uint8_t)'"'; break;
            default: c = (r == 0 && total - k == 2u) ? (uint8_t)',' : (uint8_t)'\n'; break;  // SEP
        }
        dst[k] = c;
    }
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

static void build_template(int format, Template &t) {
    const char **lits = format == KXPU_FMT_YAML ? h_yaml_lits : h_json_lits;
    uint16_t off[6];
    uint32_t acc = 0;
    for (int k = 0; k < 6; k++) { off[k] = (uint16_t)acc; acc += (uint32_t)strlen(lits[k]); }
    auto L = [&](int k) { Seg s; s.kind = LIT; s.off = off[k]; s.len = (uint16_t)strlen(lits[k]); return s; };
    auto V = [&](SegKind k) { Seg s; s.kind = k; s.off = 0; s.len = 0; return s; };
    int n = 0;
    t.segs[n++] = L(0); t.segs[n++] = V(IDX); t.segs[n++] = L(1);
    if (format == KXPU_FMT_YAML) t.segs[n++] = V(QUOTE);
    t.segs[n++] = V(BDF);
    if (format == KXPU_FMT_YAML) t.segs[n++] = V(QUOTE);
    t.segs[n++] = L(2); t.segs[n++] = V(GRP); t.segs[n++] = L(3); t.segs[n++] = V(IDX);
    t.segs[n++] = L(4); t.segs[n++] = V(GRP); t.segs[n++] = L(5);
    if (format == KXPU_FMT_JSON) t.segs[n++] = V(SEP);
    t.nsegs = n;
    t.lit_total = acc;
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
        if (cap < *len) return KXPU_E_NOSPACE;
        memcpy(out, doc, *len);
        return KXPU_OK;
    }
    const char *head = format == KXPU_FMT_YAML ? h_yaml_head : h_json_head;
    const char *tail = format == KXPU_FMT_YAML ? "" : h_json_tail;
    const size_t hl = strlen(head), tl = strlen(tail);
    const uint32_t N = (uint32_t)n;
    const size_t np = kxscan::scratch_items(n + 1);
    uint8_t *b = nullptr;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) / 256 * 256; return o; };
    size_t o_devs = take(n * sizeof(kxpu_cdidev)), o_lens = take((n + 1) * 4), o_offs = take((n + 1) * 8),
           o_part = take((np + 2) * 8), o_flags = take(16);
    KX_CUDA(ctx, cudaMallocAsync((void **)&b, off, ctx->stream));
    cudaMemcpyAsync(b + o_devs, devs, n * sizeof(kxpu_cdidev), cudaMemcpyHostToDevice, ctx->stream);
    cudaMemsetAsync(b + o_flags, 0, 16, ctx->stream);
    EmitParams E;
    E.devs = (const kxpu_cdidev *)(b + o_devs); E.n = N; E.format = format;
    build_template(format, E.tpl);
    E.lens = (uint32_t *)(b + o_lens); E.offs = (const unsigned long long *)(b + o_offs);
    E.out = nullptr; E.head_len = hl; E.flags = (uint32_t *)(b + o_flags);
    unsigned long long *d_part = (unsigned long long *)(b + o_part);
    if (ctx->stage_timing) cudaEventRecord(ctx->ev[2 * KXPU_T_EMIT], ctx->stream);
    k_cdi_len<<<(N + 1 + 255) / 256, 256, 0, ctx->stream>>>(E);
    ctx->launches++;
    kxscan::exclusive_scan<unsigned long long>(ctx, E.lens, n + 1, (unsigned long long *)(b + o_offs), d_part, nullptr);
    unsigned long long h_total = 0;
    uint32_t h_flag = 0;
    cudaMemcpyAsync(&h_total, b + o_offs + n * 8, 8, cudaMemcpyDeviceToHost, ctx->stream);
    cudaMemcpyAsync(&h_flag, b + o_flags, 4, cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    int32_t rc = KXPU_OK;
    if (e != cudaSuccess) { KX_SET_ERR(ctx, "cdi_emit sizing failed: %s", cudaGetErrorString(e)); rc = KXPU_E_CUDA; }
    else if (h_flag) { KX_SET_ERR(ctx, "cdi_emit: bdf outside [0-9a-f:.]"); rc = KXPU_E_UNSUPPORTED; }
    const size_t total = hl + (size_t)h_total + tl;
    if (rc == KXPU_OK) {
        *len = total;
        if (cap < total || !out) rc = KXPU_E_NOSPACE;
    }
    if (rc == KXPU_OK) {
        uint8_t *d_out = nullptr;
        e = cudaMallocAsync((void **)&d_out, total, ctx->stream);
        if (e == cudaSuccess) {
            E.out = d_out;
            cudaMemcpyAsync(d_out, head, hl, cudaMemcpyHostToDevice, ctx->stream);
            if (tl) cudaMemcpyAsync(d_out + hl + h_total, tail, tl, cudaMemcpyHostToDevice, ctx->stream);
            k_cdi_write<<<(N + EMIT_WARPS - 1) / EMIT_WARPS, EMIT_WARPS * 32, 0, ctx->stream>>>(E);
            ctx->launches++;
            if (ctx->stage_timing) { cudaEventRecord(ctx->ev[2 * KXPU_T_EMIT + 1], ctx->stream); ctx->ev_used[KXPU_T_EMIT] = true; }
            cudaMemcpyAsync(out, d_out, total, cudaMemcpyDeviceToHost, ctx->stream);
            cudaFreeAsync(d_out, ctx->stream);
            e = cudaStreamSynchronize(ctx->stream);
        }
        if (e != cudaSuccess) { KX_SET_ERR(ctx, "cdi_emit write failed: %s", cudaGetErrorString(e)); rc = KXPU_E_CUDA; }
    }
    cudaFreeAsync(b, ctx->stream);
    return rc;
}

// shared driver of the two "thread per item" emitters
template <typename LenK, typename WriteK>
static int32_t emit_items(kxpu_ctx *ctx, size_t n, size_t in_bytes, const void *h_in, const uint8_t *h_in2, uint8_t *out,
                          size_t cap, uint32_t *offsets, size_t *need, LenK lenk, WriteK writek) {
    const uint32_t N = (uint32_t)n;
    const size_t np = kxscan::scratch_items(n + 1);
    uint8_t *b = nullptr;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) / 256 * 256; return o; };
    size_t o_in = take(in_bytes), o_in2 = take(h_in2 ? n : 16), o_lens = take((n + 1) * 4), o_offs = take((n + 1) * 4),
           o_part = take((np + 2) * 8);
    KX_CUDA(ctx, cudaMallocAsync((void **)&b, off, ctx->stream));
    cudaMemcpyAsync(b + o_in, h_in, in_bytes, cudaMemcpyHostToDevice, ctx->stream);
    if (h_in2) cudaMemcpyAsync(b + o_in2, h_in2, n, cudaMemcpyHostToDevice, ctx->stream);
    uint32_t *d_lens = (uint32_t *)(b + o_lens), *d_offs = (uint32_t *)(b + o_offs);
    lenk(b + o_in, h_in2 ? b + o_in2 : nullptr, N, d_lens);
    ctx->launches++;
    kxscan::exclusive_scan<uint32_t>(ctx, d_lens, n + 1, d_offs, (unsigned long long *)(b + o_part), nullptr);
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
        e = cudaMallocAsync((void **)&d_out, total, ctx->stream);
        if (e == cudaSuccess) {
            writek(b + o_in, h_in2 ? b + o_in2 : nullptr, N, d_offs, d_out);
            ctx->launches++;
            cudaMemcpyAsync(out, d_out, total, cudaMemcpyDeviceToHost, ctx->stream);
            cudaFreeAsync(d_out, ctx->stream);
            e = cudaStreamSynchronize(ctx->stream);
        }
        if (e != cudaSuccess) { KX_SET_ERR(ctx, "emit write failed: %s", cudaGetErrorString(e)); rc = KXPU_E_CUDA; }
    }
    cudaFreeAsync(b, ctx->stream);
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
