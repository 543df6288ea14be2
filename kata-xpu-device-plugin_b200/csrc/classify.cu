// classify.cu -- K5: createIommuDeviceMap + device-list build over a flat record table.
//
// Reference: pkg/device_plugin/device_plugin.go:126-180 (walk, filter, group, index) and
// :91-98 (per-device-id group lists).  The sequential walk is restated as data-parallel
// primitives whose results equal the walk's:
//   candidate(i)  = !dir && vendor=="10de" && driver=="vfio-pci" && links readable   (:137-161)
//   gfirst[g]     = min{ i : candidate(i), group(i)=g, device file readable }        (:162-170:
//                   a group only comes into existence at a record whose device read works)
//   accept(i)     = candidate(i) && gfirst[group(i)] <= i                            (:171-175)
//   busIndex(i)   = #accepted before i                       -> exclusive scan
//   group ordinal = rank of gfirst[g] among group-first records -> exclusive scan
//   iommuMap CSR  = accepted records stably sorted by group ordinal -> LSD radix sort
//   deviceMap     = groups keyed by the device id of their first member, ids ordered by
//                   first appearance; CSR by a second stable sort.
#include <utility>

#include "common.cuh"
#include "scan.cuh"

namespace kxclass {

constexpr uint32_t EMPTY32 = 0xFFFFFFFFu;
constexpr unsigned long long EMPTY64 = 0xFFFFFFFFFFFFFFFFull;

struct Work {
    const kxpu_devrec *recs;
    uint32_t n;
    // group hash (keys = iommu group)
    uint32_t *gkeys, *gfirst, *gord;  // [gcap]
    uint32_t gcap, gshift;
    // devid hash (keys = packed id string)
    unsigned long long *dkeys;
    uint32_t *dfirst, *dord;  // [gcap]
    // per record
    uint32_t *gslot;   // [n] slot of the record's group (EMPTY32: not a candidate)
    uint32_t *dslot;   // [n] slot of the record's devid (group-first records only)
    unsigned long long *devid;  // [n]
    uint32_t *f_acc, *f_gf, *f_df;        // [n+1] 0/1 flags (last = 0 so the scan yields totals)
    uint32_t *s_acc, *s_gf, *s_df;        // [n+1] exclusive scans
    uint32_t *flags;   // [4] 0: unsupported input
    // sort buffers
    uint32_t *ak, *av, *ak2, *av2;  // [n] members sort
    uint32_t *bk, *bv, *bk2, *bv2;  // [n] dev_groups sort
    // outputs (device)
    uint32_t *accept_index, *group_ids, *group_off, *group_members, *dev_off, *dev_groups;
    unsigned long long *dev_ids;
};

// readIDFromFileFunc (device_plugin.go:183-191): data[2:] with '\n' trimmed at both ends.
// Returns false when the file is shorter than 2 bytes (the reference would panic) or longer
// than the 8 bytes the record carries.
__device__ __forceinline__ bool read_id(const uint8_t *txt, uint32_t flen, unsigned long long &id, uint32_t &len) {
    id = 0; len = 0;
    if (flen < 2u || flen > 8u) return false;
    int a = 2, b = (int)flen;
    while (a < b && txt[a] == (uint8_t)'\n') a++;
    while (b > a && txt[b - 1] == (uint8_t)'\n') b--;
    unsigned long long v = 0;
    for (int k = a; k < b; k++) v |= (unsigned long long)txt[k] << (8 * (k - a));
    id = v; len = (uint32_t)(b - a);
    return true;
}

__device__ __forceinline__ uint32_t hash32(uint32_t k) { return k * 0x9E3779B1u; }
__device__ __forceinline__ uint32_t hash64(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33;
    return (uint32_t)k * 0x9E3779B1u;
}

__device__ __forceinline__ uint32_t ginsert(const Work &W, uint32_t key) {
    uint32_t slot = hash32(key) >> W.gshift;
    for (;;) {
        uint32_t k = W.gkeys[slot];
        if (k == key) return slot;
        if (k == EMPTY32) {
            uint32_t old = atomicCAS(&W.gkeys[slot], EMPTY32, key);
            if (old == EMPTY32 || old == key) return slot;
        }
        slot = (slot + 1) & (W.gcap - 1);
    }
}
__device__ __forceinline__ uint32_t dinsert(const Work &W, unsigned long long key) {
    uint32_t slot = hash64(key) >> W.gshift;
    for (;;) {
        unsigned long long k = W.dkeys[slot];
        if (k == key) return slot;
        if (k == EMPTY64) {
            unsigned long long old = atomicCAS(&W.dkeys[slot], EMPTY64, key);
            if (old == EMPTY64 || old == key) return slot;
        }
        slot = (slot + 1) & (W.gcap - 1);
    }
}

// pass 1: candidates, group table, gfirst
__global__ void __launch_bounds__(256) k_candidates(const Work W) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W.n) return;
    // one 64-byte record per thread: four 16-byte vector loads
    const uint4 *rp = reinterpret_cast<const uint4 *>(W.recs + i);
    uint4 q1 = rp[1], q2 = rp[2], q3 = rp[3];  // rp[0] is the bdf: not needed to classify
    const uint8_t *vtxt = reinterpret_cast<const uint8_t *>(&q1);      // vendor_txt[8], device_txt[8]
    const uint8_t *dtxt = vtxt + 8;
    const unsigned long long drv0 = ((unsigned long long)q2.y << 32) | q2.x;  // driver[0..8)
    const uint32_t drv8 = q2.z & 0xffu;
    const uint32_t group = q3.x;
    const uint32_t vlen = q3.y & 0xffu, dlen = (q3.y >> 8) & 0xffu, fl = (q3.y >> 16) & 0xffu;
    unsigned long long vid, did;
    uint32_t vl, dl;
    bool vok = read_id(vtxt, vlen, vid, vl);
    bool cand = !(fl & KXPU_REC_IS_DIR) && !(fl & KXPU_REC_VENDOR_ERR) && vok && vl == 4u &&
                vid == 0x65643031ull /* "10de" */ && !(fl & KXPU_REC_DRIVER_ERR) &&
                drv0 == 0x6963702d6f696676ull /* "vfio-pci" */ && drv8 == 0u && !(fl & KXPU_REC_IOMMU_ERR);
    bool dok = !(fl & KXPU_REC_DEVICE_ERR) && read_id(dtxt, dlen, did, dl);
    if (!(fl & (KXPU_REC_IS_DIR | KXPU_REC_VENDOR_ERR)) && vlen > 8u) W.flags[0] = 1u;
    if (cand && (group == EMPTY32 || (dok && did == EMPTY64) || (!(fl & KXPU_REC_DEVICE_ERR) && dlen > 8u)))
        W.flags[0] = 1u;  // outside the supported domain
    uint32_t slot = EMPTY32;
    if (cand) {
        slot = ginsert(W, group);
        if (dok) atomicMin(&W.gfirst[slot], i);
    }
    W.gslot[i] = slot;
    W.devid[i] = dok ? did : EMPTY64;
}

// pass 2: accept flags, group-first flags, devid table
__global__ void __launch_bounds__(256) k_accept(const Work W) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > W.n) return;
    if (i == W.n) { W.f_acc[i] = 0; W.f_gf[i] = 0; return; }
    uint32_t slot = W.gslot[i];
    uint32_t acc = 0, gf = 0;
    if (slot != EMPTY32) {
        uint32_t first = W.gfirst[slot];
        acc = first <= i;
        gf = first == i;
    }
    W.f_acc[i] = acc;
    W.f_gf[i] = gf;
    uint32_t ds = EMPTY32;
    if (gf) {
        ds = dinsert(W, W.devid[i]);
        atomicMin(&W.dfirst[ds], i);
    }
    W.dslot[i] = ds;
}

// pass 3: busIndex, group ordinals, dev-first flags
__global__ void __launch_bounds__(256) k_groups(const Work W) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > W.n) return;
    if (i == W.n) { W.f_df[i] = 0; return; }
    W.accept_index[i] = W.f_acc[i] ? W.s_acc[i] : KXPU_REJECTED;
    uint32_t df = 0;
    if (W.f_gf[i]) {
        uint32_t ord = W.s_gf[i];
        W.group_ids[ord] = W.gkeys[W.gslot[i]];
        W.gord[W.gslot[i]] = ord;
        df = W.dfirst[W.dslot[i]] == i;
    }
    W.f_df[i] = df;
}

// pass 4: dev ordinals
__global__ void __launch_bounds__(256) k_devids(const Work W) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W.n) return;
    if (W.f_df[i]) {
        uint32_t ord = W.s_df[i];
        W.dev_ids[ord] = W.devid[i];
        W.dord[W.dslot[i]] = ord;
    }
}

// pass 5: sort inputs.  members: (group ordinal, record) at busIndex; groups: (dev ordinal, group id) at group ordinal
__global__ void __launch_bounds__(256) k_pairs(const Work W) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W.n) return;
    if (W.f_acc[i]) {
        uint32_t b = W.s_acc[i];
        W.ak[b] = W.gord[W.gslot[i]];
        W.av[b] = i;
    }
    if (W.f_gf[i]) {
        uint32_t o = W.s_gf[i];
        W.bk[o] = W.dord[W.dslot[i]];
        W.bv[o] = W.gkeys[W.gslot[i]];
    }
}

// ---------------------------------------------------------------- stable LSD radix sort
constexpr int RS_THREADS = 256;
constexpr int RS_ROUNDS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ROUNDS;

__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const uint32_t *__restrict__ keys, const uint32_t *d_count,
                                                        uint32_t shift, uint32_t nblocks, uint32_t *__restrict__ hist) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t n = *d_count;
    const uint32_t base = blockIdx.x * RS_TILE;
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; r++) {
        uint32_t i = base + r * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];  // digit-major so one scan orders it
}

__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const uint32_t *__restrict__ keys,
                                                           const uint32_t *__restrict__ vals, const uint32_t *d_count,
                                                           uint32_t shift, uint32_t nblocks,
                                                           const uint32_t *__restrict__ hist_scan,
                                                           uint32_t *__restrict__ okeys, uint32_t *__restrict__ ovals) {
    __shared__ uint32_t dbase[256];
    __shared__ uint32_t wcnt[RS_THREADS / 32][256];
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    dbase[threadIdx.x] = hist_scan[threadIdx.x * nblocks + blockIdx.x];
    const uint32_t n = *d_count;
    const uint32_t base = blockIdx.x * RS_TILE;
    if (base >= n) return;
    for (int r = 0; r < RS_ROUNDS; r++) {
#pragma unroll
        for (int k = 0; k < RS_THREADS / 32; k++) wcnt[k][threadIdx.x] = 0;
        __syncthreads();
        const uint32_t i = base + r * RS_THREADS + threadIdx.x;
        const bool act = i < n;
        uint32_t key = 0, val = 0, d = 256u;  // 256 = inactive, never matches a digit
        if (act) { key = keys[i]; val = vals[i]; d = (key >> shift) & 255u; }
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const uint32_t rank = (uint32_t)__popc(peers & ((1u << lane) - 1u));
        if (act && rank == 0) wcnt[w][d] = (uint32_t)__popc(peers);
        __syncthreads();
        if (act) {
            uint32_t before = 0;
            for (uint32_t k = 0; k < w; k++) before += wcnt[k][d];
            uint32_t pos = dbase[d] + before + rank;
            okeys[pos] = key;
            ovals[pos] = val;
        }
        __syncthreads();
        uint32_t tot = 0;
#pragma unroll
        for (int k = 0; k < RS_THREADS / 32; k++) tot += wcnt[k][threadIdx.x];
        dbase[threadIdx.x] += tot;
        __syncthreads();
    }
}

// off[key[j]] = j at every run start; off[n_ord] = count
__global__ void __launch_bounds__(256) k_bounds(const uint32_t *__restrict__ keys, const uint32_t *d_count,
                                                const uint32_t *d_nord, uint32_t *__restrict__ off) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = *d_count;
    if (j == 0) off[*d_nord] = n;
    if (j >= n) return;
    if (j == 0 || keys[j] != keys[j - 1]) off[keys[j]] = j;
}

}  // namespace kxclass

using namespace kxclass;

static uint32_t bits_for(uint32_t n) {
    uint32_t b = 1;
    while (b < 32 && (1ull << b) < (unsigned long long)n + 1) b++;
    return b;
}

// sorts (k,v) pairs of *d_count items (<= n_max) by the low `bits` bits of k; result ends in (k,v)
static void radix_sort_pairs(kxpu_ctx *ctx, uint32_t *&k, uint32_t *&v, uint32_t *&k2, uint32_t *&v2, uint32_t n_max,
                             const uint32_t *d_count, uint32_t bits, uint32_t *d_hist, uint32_t *d_hist_scan,
                             unsigned long long *d_part) {
    const uint32_t nblocks = (n_max + RS_TILE - 1) / RS_TILE;
    for (uint32_t shift = 0; shift < bits; shift += 8) {
        k_rs_hist<<<nblocks, RS_THREADS, 0, ctx->stream>>>(k, d_count, shift, nblocks, d_hist);
        ctx->launches++;
        kxscan::exclusive_scan<uint32_t>(ctx, d_hist, (size_t)256 * nblocks, d_hist_scan, d_part, nullptr);
        k_rs_scatter<<<nblocks, RS_THREADS, 0, ctx->stream>>>(k, v, d_count, shift, nblocks, d_hist_scan, k2, v2);
        ctx->launches++;
        std::swap(k, k2);
        std::swap(v, v2);
    }
}

extern "C" int32_t kxpu_classify(kxpu_ctx *ctx, const kxpu_devrec *recs, size_t n, kxpu_classify_out *out) {
    if (!ctx || !out || (n && !recs)) return KXPU_E_INVALID;
    if (n >= 0x7FFFFFFFull) return KXPU_E_UNSUPPORTED;
    std::lock_guard<std::mutex> guard(ctx->mu);
    cudaSetDevice(ctx->device);
    kx_clear_timings(ctx);
    out->n_accepted = out->n_groups = out->n_devids = 0;
    if (n == 0) {
        if (out->group_off) out->group_off[0] = 0;
        if (out->dev_off) out->dev_off[0] = 0;
        return KXPU_OK;
    }
    if (!out->accept_index || !out->group_ids || !out->group_off || !out->group_members || !out->dev_ids ||
        !out->dev_off || !out->dev_groups)
        return KXPU_E_INVALID;

    const uint32_t N = (uint32_t)n;
    uint32_t gcap = 1024;
    while (gcap < 2 * N) gcap <<= 1;
    uint32_t lg = 0;
    while ((1u << lg) < gcap) lg++;
    const uint32_t nblocks_rs = (N + RS_TILE - 1) / RS_TILE;

    // one arena; [ff-region | zero-region | rest]
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) / 256 * 256; return o; };
    size_t o_gkeys = take((size_t)gcap * 4), o_gfirst = take((size_t)gcap * 4), o_dkeys = take((size_t)gcap * 8),
           o_dfirst = take((size_t)gcap * 4);
    size_t ff_bytes = off;
    size_t o_flags = take(16);
    size_t zero_bytes = off - ff_bytes;
    size_t o_gord = take((size_t)gcap * 4), o_dord = take((size_t)gcap * 4);
    size_t o_recs = take(n * sizeof(kxpu_devrec));
    size_t o_gslot = take(n * 4), o_dslot = take(n * 4), o_devid = take(n * 8);
    size_t o_facc = take((n + 1) * 4), o_fgf = take((n + 1) * 4), o_fdf = take((n + 1) * 4);
    size_t o_sacc = take((n + 1) * 4), o_sgf = take((n + 1) * 4), o_sdf = take((n + 1) * 4);
    size_t o_ak = take(n * 4), o_av = take(n * 4), o_ak2 = take(n * 4), o_av2 = take(n * 4);
    size_t o_bk = take(n * 4), o_bv = take(n * 4), o_bk2 = take(n * 4), o_bv2 = take(n * 4);
    size_t o_hist = take((size_t)256 * nblocks_rs * 4), o_hscan = take((size_t)256 * nblocks_rs * 4);
    size_t o_part = take((kxscan::scratch_items((size_t)256 * nblocks_rs + n + 2) + 2) * 8);
    size_t o_acc_idx = take(n * 4), o_gids = take(n * 4), o_goff = take((n + 1) * 4), o_gmem = take(n * 4);
    size_t o_dids = take(n * 8), o_doff = take((n + 1) * 4), o_dgrp = take(n * 4);
    uint8_t *b = nullptr;
    KX_CUDA(ctx, cudaMallocAsync((void **)&b, off, ctx->stream));
    cudaMemsetAsync(b, 0xff, ff_bytes, ctx->stream);
    cudaMemsetAsync(b + ff_bytes, 0, zero_bytes, ctx->stream);

    Work W;
    W.recs = (const kxpu_devrec *)(b + o_recs); W.n = N;
    W.gkeys = (uint32_t *)(b + o_gkeys); W.gfirst = (uint32_t *)(b + o_gfirst); W.gord = (uint32_t *)(b + o_gord);
    W.gcap = gcap; W.gshift = 32 - lg;
    W.dkeys = (unsigned long long *)(b + o_dkeys); W.dfirst = (uint32_t *)(b + o_dfirst); W.dord = (uint32_t *)(b + o_dord);
    W.gslot = (uint32_t *)(b + o_gslot); W.dslot = (uint32_t *)(b + o_dslot); W.devid = (unsigned long long *)(b + o_devid);
    W.f_acc = (uint32_t *)(b + o_facc); W.f_gf = (uint32_t *)(b + o_fgf); W.f_df = (uint32_t *)(b + o_fdf);
    W.s_acc = (uint32_t *)(b + o_sacc); W.s_gf = (uint32_t *)(b + o_sgf); W.s_df = (uint32_t *)(b + o_sdf);
    W.flags = (uint32_t *)(b + o_flags);
    W.ak = (uint32_t *)(b + o_ak); W.av = (uint32_t *)(b + o_av); W.ak2 = (uint32_t *)(b + o_ak2); W.av2 = (uint32_t *)(b + o_av2);
    W.bk = (uint32_t *)(b + o_bk); W.bv = (uint32_t *)(b + o_bv); W.bk2 = (uint32_t *)(b + o_bk2); W.bv2 = (uint32_t *)(b + o_bv2);
    W.accept_index = (uint32_t *)(b + o_acc_idx); W.group_ids = (uint32_t *)(b + o_gids);
    W.group_off = (uint32_t *)(b + o_goff); W.group_members = (uint32_t *)(b + o_gmem);
    W.dev_ids = (unsigned long long *)(b + o_dids); W.dev_off = (uint32_t *)(b + o_doff); W.dev_groups = (uint32_t *)(b + o_dgrp);
    uint32_t *d_hist = (uint32_t *)(b + o_hist), *d_hscan = (uint32_t *)(b + o_hscan);
    unsigned long long *d_part = (unsigned long long *)(b + o_part);

    cudaMemcpyAsync(b + o_recs, recs, n * sizeof(kxpu_devrec), cudaMemcpyHostToDevice, ctx->stream);
    const unsigned g = (N + 255) / 256, g1 = (N + 1 + 255) / 256;
    {
        KxTimer tm(ctx, KXPU_T_CLASSIFY);
        k_candidates<<<g, 256, 0, ctx->stream>>>(W);
        k_accept<<<g1, 256, 0, ctx->stream>>>(W);
        ctx->launches += 2;
        kxscan::exclusive_scan<uint32_t>(ctx, W.f_acc, n + 1, W.s_acc, d_part, nullptr);
        kxscan::exclusive_scan<uint32_t>(ctx, W.f_gf, n + 1, W.s_gf, d_part, nullptr);
        k_groups<<<g1, 256, 0, ctx->stream>>>(W);
        ctx->launches++;
        kxscan::exclusive_scan<uint32_t>(ctx, W.f_df, n + 1, W.s_df, d_part, nullptr);
        k_devids<<<g, 256, 0, ctx->stream>>>(W);
        k_pairs<<<g, 256, 0, ctx->stream>>>(W);
        ctx->launches += 2;
        // totals live at index n of the scans
        const uint32_t *d_nacc = W.s_acc + n, *d_ngrp = W.s_gf + n, *d_ndev = W.s_df + n;
        const uint32_t bits = bits_for(N);
        radix_sort_pairs(ctx, W.ak, W.av, W.ak2, W.av2, N, d_nacc, bits, d_hist, d_hscan, d_part);
        radix_sort_pairs(ctx, W.bk, W.bv, W.bk2, W.bv2, N, d_ngrp, bits, d_hist, d_hscan, d_part);
        k_bounds<<<g, 256, 0, ctx->stream>>>(W.ak, d_nacc, d_ngrp, W.group_off);
        k_bounds<<<g, 256, 0, ctx->stream>>>(W.bk, d_ngrp, d_ndev, W.dev_off);
        ctx->launches += 2;
    }
    // results: sorted values are the CSR payloads
    uint32_t *h = ctx->h_ctl;
    cudaMemcpyAsync(&h[0], W.s_acc + n, 4, cudaMemcpyDeviceToHost, ctx->stream);
    cudaMemcpyAsync(&h[1], W.s_gf + n, 4, cudaMemcpyDeviceToHost, ctx->stream);
    cudaMemcpyAsync(&h[2], W.s_df + n, 4, cudaMemcpyDeviceToHost, ctx->stream);
    cudaMemcpyAsync(&h[3], W.flags, 4, cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    int32_t rc = KXPU_OK;
    if (e != cudaSuccess) { KX_SET_ERR(ctx, "classify failed: %s", cudaGetErrorString(e)); rc = KXPU_E_CUDA; }
    else if (h[3]) { KX_SET_ERR(ctx, "classify: record outside the supported domain (group 0xffffffff or id file > 8 bytes)"); rc = KXPU_E_UNSUPPORTED; }
    if (rc == KXPU_OK) {
        const uint32_t na = h[0], ng = h[1], nd = h[2];
        out->n_accepted = na; out->n_groups = ng; out->n_devids = nd;
        cudaMemcpyAsync(out->accept_index, W.accept_index, n * 4, cudaMemcpyDeviceToHost, ctx->stream);
        cudaMemcpyAsync(out->group_ids, W.group_ids, (size_t)ng * 4, cudaMemcpyDeviceToHost, ctx->stream);
        cudaMemcpyAsync(out->group_off, W.group_off, ((size_t)ng + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream);
        cudaMemcpyAsync(out->group_members, W.av, (size_t)na * 4, cudaMemcpyDeviceToHost, ctx->stream);
        cudaMemcpyAsync(out->dev_ids, W.dev_ids, (size_t)nd * 8, cudaMemcpyDeviceToHost, ctx->stream);
        cudaMemcpyAsync(out->dev_off, W.dev_off, ((size_t)nd + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream);
        cudaMemcpyAsync(out->dev_groups, W.bv, (size_t)ng * 4, cudaMemcpyDeviceToHost, ctx->stream);
        e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { KX_SET_ERR(ctx, "classify D2H failed: %s", cudaGetErrorString(e)); rc = KXPU_E_CUDA; }
        if (ng == 0) out->group_off[0] = 0;
        if (nd == 0) out->dev_off[0] = 0;
    }
    cudaFreeAsync(b, ctx->stream);
    return rc;
}
