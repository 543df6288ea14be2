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
// Launches: reset | candidates | accept + both scans (single pass, decoupled look-back) | per-group device ids | device-id
// first-seen scan over the groups | sort pairs + all digit histograms | <= 4 radix passes, each ONE
// kernel for both sorts (per-tile ranking + per-digit look-back, "onesweep") | CSR boundaries.
#include <algorithm>
#include <utility>

#include "common.cuh"
#include "scan.cuh"

namespace kxclass {

constexpr uint32_t EMPTY32 = 0xFFFFFFFFu;
constexpr unsigned long long EMPTY64 = 0xFFFFFFFFFFFFFFFFull;

struct __align__(16) GSlot { uint32_t key, first, ord, pad; };                // iommu group -> first good record, ordinal
struct __align__(16) DSlot { unsigned long long key; uint32_t first, ord; };  // device id string -> first group-first record, ordinal

constexpr int C_THREADS = 256;
constexpr int C_ITEMS = 8;
constexpr int C_TILE = C_THREADS * C_ITEMS;  // records per CTA of the scan kernels

// totals[] (device): 0 accepted, 1 groups, 2 device ids, 3 unsupported-input flag
struct Work {
    const kxpu_devrec *recs;
    uint32_t n;
    GSlot *gtab;
    DSlot *dtab;
    uint32_t gcap, gshift, dcap, dshift;
    uint32_t *gslot;      // [n] slot of the record's group (EMPTY32: not a candidate)
    uint32_t *grp_rec;    // [n_groups] first record of group ordinal o
    uint32_t *grp_dslot;  // [n_groups] device-id slot of group ordinal o
    uint32_t *totals;
    unsigned long long *st_acc, *st_gf, *st_df;  // look-back status words
    uint32_t ep_acc, ep_gf, ep_df;
    // sort buffers: members (group ordinal, record) / groups (device ordinal, group id)
    uint32_t *ak, *av, *bk, *bv;
    uint32_t *ghist;  // [2 sorts][4 passes][256]
    // outputs (device)
    uint32_t *accept_index, *group_ids, *group_off, *dev_off;
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
        uint32_t k = __ldcg(&W.gtab[slot].key);
        if (k == key) return slot;
        if (k == EMPTY32) {
            uint32_t old = atomicCAS(&W.gtab[slot].key, EMPTY32, key);
            if (old == EMPTY32 || old == key) return slot;
        }
        slot = (slot + 1) & (W.gcap - 1);
    }
}
// the device-id table starts small (real hosts see a few ids; a 4-hex id space holds 65 536): a probe
// run of 512 means it is too small -- flag it (totals[3] bit 1), the host runs again with dcap = gcap
__device__ __forceinline__ uint32_t dinsert(const Work &W, unsigned long long key) {
    uint32_t slot = hash64(key) >> W.dshift;
    for (uint32_t step = 0; step < 512u; step++) {
        unsigned long long k = __ldcg(&W.dtab[slot].key);
        if (k == key) return slot;
        if (k == EMPTY64) {
            unsigned long long old = atomicCAS(&W.dtab[slot].key, EMPTY64, key);
            if (old == EMPTY64 || old == key) return slot;
        }
        slot = (slot + 1) & (W.dcap - 1);
    }
    atomicOr(&W.totals[3], 2u);
    return 0u;
}

// pass 1: candidates, group table, gfirst
__global__ void __launch_bounds__(256) k_candidates(const Work W) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W.n) return;
    // one 64-byte record per thread: three 16-byte vector loads (the bdf is not needed to classify)
    const uint4 *rp = reinterpret_cast<const uint4 *>(W.recs + i);
    uint4 q1 = rp[1], q2 = rp[2], q3 = rp[3];
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
    if (!(fl & (KXPU_REC_IS_DIR | KXPU_REC_VENDOR_ERR)) && vlen > 8u) atomicOr(&W.totals[3], 1u);
    if (cand && (group == EMPTY32 || (dok && did == EMPTY64) || (!(fl & KXPU_REC_DEVICE_ERR) && dlen > 8u)))
        atomicOr(&W.totals[3], 1u);  // outside the supported domain
    uint32_t slot = EMPTY32;
    if (cand) {
        slot = ginsert(W, group);
        if (dok && i < __ldcg(&W.gtab[slot].first)) atomicMin(&W.gtab[slot].first, i);
    }
    W.gslot[i] = slot;
}

// pass 2: accept / group-first flags of a 2048-record tile, both exclusive scans in the same kernel
// (two look-backs, warp 0 and warp 1), busIndex out, group ordinals out, device-id table insert.
__global__ void __launch_bounds__(C_THREADS) k_accept_scan(const Work W) {
    __shared__ uint32_t wsum[C_THREADS / 32];
    __shared__ uint32_t s_tot;
    __shared__ unsigned long long s_excl[2];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
    const uint32_t base = blockIdx.x * C_TILE + tid * C_ITEMS;  // blocked: 16 consecutive records per thread
    uint32_t slot[C_ITEMS];
    if (base + C_ITEMS <= W.n) {
#pragma unroll
        for (int k = 0; k < C_ITEMS; k += 4) {
            const uint4 q = *reinterpret_cast<const uint4 *>(W.gslot + base + k);
            slot[k] = q.x; slot[k + 1] = q.y; slot[k + 2] = q.z; slot[k + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < C_ITEMS; k++) slot[k] = base + k < W.n ? W.gslot[base + k] : EMPTY32;
    }
    uint32_t accm = 0, gfm = 0;  // bit k: record base + k accepted / first of its group
    uint32_t first[C_ITEMS];
#pragma unroll
    for (int k = 0; k < C_ITEMS; k++) first[k] = W.gtab[slot[k] != EMPTY32 ? slot[k] : 0u].first;  // all loads in flight at once
#pragma unroll
    for (int k = 0; k < C_ITEMS; k++) {
        if (slot[k] != EMPTY32) {
            if (first[k] <= base + k) accm |= 1u << k;
            if (first[k] == base + k) gfm |= 1u << k;
        }
    }
    const uint32_t packed = (uint32_t)__popc(accm) | ((uint32_t)__popc(gfm) << 16);  // <= 2048 each per tile
    const uint32_t incl = kxscan::warp_incl(packed);
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    if (w == 0) {
        const uint32_t x = lane < C_THREADS / 32 ? wsum[lane] : 0u;
        const uint32_t xi = kxscan::warp_incl(x);
        if (lane < C_THREADS / 32) wsum[lane] = xi - x;
        if (lane == C_THREADS / 32 - 1) s_tot = xi;
    }
    __syncthreads();
    const uint32_t ex = wsum[w] + incl - packed;
    const uint32_t tot = s_tot;
    if (w < 2) {
        const unsigned long long agg = w == 0 ? (tot & 0xffffu) : (tot >> 16);
        const unsigned long long e = kxscan::lookback(w == 0 ? W.st_acc : W.st_gf, blockIdx.x, agg, w == 0 ? W.ep_acc : W.ep_gf);
        if (lane == 0) {
            s_excl[w] = e;
            if (blockIdx.x == gridDim.x - 1) W.totals[w] = (uint32_t)(e + agg);
        }
    }
    __syncthreads();
    uint32_t racc = (uint32_t)s_excl[0] + (ex & 0xffffu), rgf = (uint32_t)s_excl[1] + (ex >> 16);
    uint32_t outv[C_ITEMS];
#pragma unroll
    for (int k = 0; k < C_ITEMS; k++) {
        outv[k] = KXPU_REJECTED;
        if ((accm >> k) & 1u) outv[k] = racc++;
        if ((gfm >> k) & 1u) {
            const uint32_t ord = rgf++;
            W.gtab[slot[k]].ord = ord;
            W.grp_rec[ord] = base + k;  // the rest of the group's bookkeeping: k_groups, one thread per group
        }
    }
    if (base + C_ITEMS <= W.n) {
#pragma unroll
        for (int k = 0; k < C_ITEMS; k += 4)
            *reinterpret_cast<uint4 *>(W.accept_index + base + k) = make_uint4(outv[k], outv[k + 1], outv[k + 2], outv[k + 3]);
    } else {
#pragma unroll
        for (int k = 0; k < C_ITEMS; k++)
            if (base + k < W.n) W.accept_index[base + k] = outv[k];
    }
}

// pass 2b: one thread per group (ordinal order): the group id and the device id of its first member (the group is
// attributed to the device id of its FIRST member, device_plugin.go:162-170) -> device-id table, first-seen minimum.
// Kept out of k_accept_scan: there the chain record read -> table insert -> minimum ran serially per record in a
// divergent loop (6 % issue utilisation); here every group is an independent thread.
__global__ void __launch_bounds__(256) k_groups(const Work W) {
    const uint32_t o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= W.totals[1]) return;
    const uint32_t i = W.grp_rec[o];
    const uint32_t *rw = reinterpret_cast<const uint32_t *>(W.recs + i);
    const uint2 dq = make_uint2(rw[6], rw[7]);  // device_txt
    const uint32_t dlen = (rw[13] >> 8) & 0xffu;
    W.group_ids[o] = rw[12];                    // iommu_group
    unsigned long long did;
    uint32_t dl;
    read_id(reinterpret_cast<const uint8_t *>(&dq), dlen, did, dl);
    const uint32_t ds = dinsert(W, did);
    // a few hot device ids own most groups: same-address atomics run at ~1 per ns, so only a group that can
    // still lower the minimum issues one
    if (i < __ldcg(&W.dtab[ds].first)) atomicMin(&W.dtab[ds].first, i);
    W.grp_dslot[o] = ds;
}

// pass 3: over the groups in ordinal order: is this the first group of its device id?  Scan -> device ordinals.
__global__ void __launch_bounds__(C_THREADS) k_devfirst_scan(const Work W) {
    __shared__ unsigned long long s_excl;
    const uint32_t ng = W.totals[1];
    const uint32_t base = blockIdx.x * C_TILE + threadIdx.x * C_ITEMS;
    uint32_t dfm = 0;
#pragma unroll
    for (int k = 0; k < C_ITEMS; k++) {
        const uint32_t o = base + k;
        if (o < ng && W.dtab[W.grp_dslot[o]].first == W.grp_rec[o]) dfm |= 1u << k;
    }
    uint32_t tot;
    const uint32_t ex = kxscan::block_excl((uint32_t)__popc(dfm), &tot);
    if (threadIdx.x < 32) {
        const unsigned long long e = kxscan::lookback(W.st_df, blockIdx.x, tot, W.ep_df);
        if (threadIdx.x == 0) {
            s_excl = e;
            if (blockIdx.x == gridDim.x - 1) W.totals[2] = (uint32_t)(e + tot);
        }
    }
    __syncthreads();
    uint32_t run = (uint32_t)s_excl + ex;
#pragma unroll
    for (int k = 0; k < C_ITEMS; k++) {
        if ((dfm >> k) & 1u) {
            DSlot &d = W.dtab[W.grp_dslot[base + k]];
            d.ord = run;
            W.dev_ids[run] = d.key;
            run++;
        }
    }
}

// pass 4: sort inputs -- members (group ordinal, record) at busIndex, groups (device ordinal, group id)
// at group ordinal -- and the digit histograms of all radix passes of both sorts.
__global__ void __launch_bounds__(256) k_pairs(const Work W, uint32_t passes) {
    __shared__ uint32_t h[2 * 4 * 256];
    for (uint32_t k = threadIdx.x; k < 2 * 4 * 256; k += 256) h[k] = 0u;
    __syncthreads();
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t ng = W.totals[1];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < W.n; i += stride) {
        const uint32_t b = W.accept_index[i];
        if (b != KXPU_REJECTED) {
            const uint32_t key = W.gtab[W.gslot[i]].ord;
            W.ak[b] = key;
            W.av[b] = i;
            for (uint32_t p = 0; p < passes; p++) atomicAdd(&h[(0 * 4 + p) * 256 + ((key >> (8 * p)) & 255u)], 1u);
        }
        if (i < ng) {
            const uint32_t key = W.dtab[W.grp_dslot[i]].ord;
            W.bk[i] = key;
            W.bv[i] = W.group_ids[i];
            for (uint32_t p = 0; p < passes; p++) atomicAdd(&h[(1 * 4 + p) * 256 + ((key >> (8 * p)) & 255u)], 1u);
        }
    }
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < 2 * 4 * 256; k += 256) {
        const uint32_t v = h[k];
        if (v) atomicAdd(&W.ghist[k], v);
    }
}

// ---------------------------------------------------------------- stable LSD radix sort, one kernel per pass
// A CTA ranks a tile of 4096 pairs: every warp owns 256 consecutive items and counts digits in its
// own shared-memory counters (rank inside the warp by __match_any_sync), the warps' counts are
// scanned per digit, the tile's count of every digit is published and the digit's offset over the
// tiles in front comes from a look-back over those status words (one digit per thread), the
// pass-wide digit bases from the histogram k_pairs made.  Both sorts run in the same launch.
constexpr int OS_WARPS = 16;
constexpr int OS_THREADS = OS_WARPS * 32;
constexpr int OS_STEPS = 8;
constexpr int OS_TILE = OS_THREADS * OS_STEPS;  // 4096

struct SortJob {
    const uint32_t *kin, *vin;
    uint32_t *kout, *vout;
    const uint32_t *count;      // items (device)
    const uint32_t *ghist;      // [256] digit totals of this pass
    unsigned long long *state;  // [tiles][256]
};
struct SweepParams {
    SortJob job[2];
    uint32_t shift, epoch;
};

__global__ void __launch_bounds__(OS_THREADS) k_onesweep(const SweepParams P) {
    __shared__ uint32_t cnt[OS_WARPS][256];
    __shared__ uint32_t tbase[256];
    __shared__ uint32_t wsum[OS_THREADS / 32];
    const SortJob J = blockIdx.y ? P.job[1] : P.job[0];  // constant indices: the parameters stay in the constant bank
    const uint32_t tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
    const uint32_t n = *J.count;
    const uint32_t tile = blockIdx.x;
    const uint32_t wbase = tile * OS_TILE + w * (OS_TILE / OS_WARPS);
#pragma unroll
    for (int k = 0; k < 256 / 32; k++) cnt[w][lane + 32 * k] = 0u;
    __syncwarp();
    uint32_t key[OS_STEPS], val[OS_STEPS], rk[OS_STEPS];
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int s = 0; s < OS_STEPS; s++) {
        const uint32_t i = wbase + s * 32u + lane;
        const bool act = i < n;
        key[s] = act ? J.kin[i] : 0u;
        val[s] = act ? J.vin[i] : 0u;
        const uint32_t d = act ? ((key[s] >> P.shift) & 255u) : 256u;  // 256: inactive lanes match each other only
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const uint32_t r = (uint32_t)__popc(peers & lt);
        uint32_t prev = 0;
        if (act) prev = cnt[w][d];
        __syncwarp();
        if (act && r == 0u) cnt[w][d] = prev + (uint32_t)__popc(peers);
        __syncwarp();
        rk[s] = prev + r;
    }
    __syncthreads();
    // per digit: exclusive scan over the warps, tile count
    uint32_t tot = 0;
    if (tid < 256) {
#pragma unroll
        for (int k = 0; k < OS_WARPS; k++) {
            const uint32_t x = cnt[k][tid];
            cnt[k][tid] = tot;
            tot += x;
        }
    }
    // pass-wide digit bases: exclusive scan of the 256 digit totals
    const uint32_t gh = tid < 256 ? J.ghist[tid] : 0u;
    const uint32_t gi = kxscan::warp_incl(gh);
    if (lane == 31) wsum[w] = gi;
    __syncthreads();
    if (w == 0) {
        const uint32_t x = lane < OS_THREADS / 32 ? wsum[lane] : 0u;
        const uint32_t xi = kxscan::warp_incl(x);
        if (lane < OS_THREADS / 32) wsum[lane] = xi - x;
    }
    __syncthreads();
    if (tid < 256) {
        const uint32_t gbase = wsum[w] + gi - gh;
        // look-back over the tiles in front, for this thread's digit
        const unsigned long long tag = (unsigned long long)(P.epoch & 0xffffffu) << kxscan::ST_EPOCH_SHIFT;
        unsigned long long *st = J.state + tid;
        *reinterpret_cast<volatile unsigned long long *>(st + (size_t)tile * 256) = tag | (tile == 0 ? kxscan::ST_PFX : kxscan::ST_AGG) | tot;
        uint32_t excl = 0;
        for (long long j = (long long)tile - 1; j >= 0;) {
            const unsigned long long v = kxscan::ld_state(st + (size_t)j * 256);
            if ((v >> kxscan::ST_EPOCH_SHIFT) != (tag >> kxscan::ST_EPOCH_SHIFT) || (v & kxscan::ST_FLAGS) == 0) continue;  // not published yet
            excl += (uint32_t)(v & kxscan::ST_VAL);
            if ((v & kxscan::ST_FLAGS) == kxscan::ST_PFX) break;
            j--;
        }
        if (tile != 0) *reinterpret_cast<volatile unsigned long long *>(st + (size_t)tile * 256) = tag | kxscan::ST_PFX | (unsigned long long)(excl + tot);
        tbase[tid] = gbase + excl;
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < OS_STEPS; s++) {
        const uint32_t i = wbase + s * 32u + lane;
        if (i < n) {
            const uint32_t d = (key[s] >> P.shift) & 255u;
            const uint32_t pos = tbase[d] + cnt[w][d] + rk[s];
            J.kout[pos] = key[s];
            J.vout[pos] = val[s];
        }
    }
}

// off[key[j]] = j at every run start; off[n_ord] = count.  blockIdx.y: members / groups
struct BoundsParams {
    const uint32_t *keys[2], *count[2], *nord[2];
    uint32_t *off[2];
};
__global__ void __launch_bounds__(256) k_bounds(const BoundsParams B) {
    const uint32_t *keys = blockIdx.y ? B.keys[1] : B.keys[0];
    const uint32_t *nord = blockIdx.y ? B.nord[1] : B.nord[0];
    uint32_t *off = blockIdx.y ? B.off[1] : B.off[0];
    const uint32_t n = blockIdx.y ? *B.count[1] : *B.count[0];
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t j0 = blockIdx.x * blockDim.x + threadIdx.x;
    if (j0 == 0) off[*nord] = n;
    for (uint32_t j = j0; j < n; j += stride) {
        const uint32_t k = keys[j];
        if (j == 0 || k != keys[j - 1]) off[k] = j;
    }
}

// 0xff over the hash tables, 0 over totals + histograms: one launch
__global__ void __launch_bounds__(256) k_reset(uint4 *ff, size_t n_ff16, uint32_t *zero, uint32_t n_zero) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t k = i; k < n_ff16; k += stride) ff[k] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    for (size_t k = i; k < n_zero; k += stride) zero[k] = 0u;
}

}  // namespace kxclass

using namespace kxclass;

static uint32_t bits_for(uint32_t n) {
    uint32_t b = 1;
    while (b < 32 && (1ull << b) < (unsigned long long)n + 1) b++;
    return b;
}

static int32_t classify_once(kxpu_ctx *ctx, const kxpu_devrec *recs, size_t n, kxpu_classify_out *out, bool small_dtab, bool *retry);

extern "C" int32_t kxpu_classify(kxpu_ctx *ctx, const kxpu_devrec *recs, size_t n, kxpu_classify_out *out) {
    if (!ctx || !out || (n && !recs)) return KXPU_E_INVALID;
    if (n >= 0x7FFFFFFFull) return KXPU_E_UNSUPPORTED;
    std::lock_guard<std::mutex> guard(ctx->mu);
    cudaSetDevice(ctx->device);
    kx_clear_timings(ctx);
    bool retry = false;
    int32_t rc = classify_once(ctx, recs, n, out, true, &retry);
    if (retry) rc = classify_once(ctx, recs, n, out, false, &retry);  // more distinct device ids than the small table holds
    return rc;
}

static int32_t classify_once(kxpu_ctx *ctx, const kxpu_devrec *recs, size_t n, kxpu_classify_out *out, bool small_dtab, bool *retry) {
    *retry = false;
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
    const uint32_t dcap = small_dtab ? std::min<uint32_t>(gcap, 1u << 17) : gcap;
    uint32_t dlg = 0;
    while ((1u << dlg) < dcap) dlg++;
    const uint32_t c_tiles = (N + C_TILE - 1) / C_TILE;
    const uint32_t s_tiles = (N + OS_TILE - 1) / OS_TILE;
    const uint32_t passes = (bits_for(N) + 7) / 8;

    // one arena; [ff-region | zero-region | rest]
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) / 256 * 256; return o; };
    const size_t o_gtab = take((size_t)gcap * sizeof(GSlot)), o_dtab = take((size_t)dcap * sizeof(DSlot));
    const size_t ff_bytes = off;
    const size_t o_totals = take(16), o_ghist = take(2 * 4 * 256 * 4);
    const size_t zero_words = (off - ff_bytes) / 4;
    const size_t o_recs = take(n * sizeof(kxpu_devrec));
    const size_t o_gslot = take(n * 4 + 64), o_grec = take(n * 4), o_gds = take(n * 4);
    const size_t o_ak = take(n * 4), o_av = take(n * 4), o_ak2 = take(n * 4), o_av2 = take(n * 4);
    const size_t o_bk = take(n * 4), o_bv = take(n * 4), o_bk2 = take(n * 4), o_bv2 = take(n * 4);
    const size_t o_acc_idx = take(n * 4 + 64), o_gids = take(n * 4), o_goff = take((n + 1) * 4);
    const size_t o_dids = take(n * 8), o_doff = take((n + 1) * 4);
    KxScratch sc(ctx);
    uint8_t *b = nullptr;
    KX_CUDA(ctx, sc.alloc((void **)&b, off));
    // look-back status words: three scans + the two sorts' per-digit words
    const size_t st_words = 3 * (size_t)c_tiles + 2 * (size_t)s_tiles * 256;
    unsigned long long *st = kx_scan_state(ctx, st_words);
    if (!st) return KXPU_E_NOMEM;

    Work W;
    memset(&W, 0, sizeof W);
    W.recs = (const kxpu_devrec *)(b + o_recs); W.n = N;
    W.gtab = (GSlot *)(b + o_gtab); W.dtab = (DSlot *)(b + o_dtab); W.gcap = gcap; W.gshift = 32 - lg; W.dcap = dcap; W.dshift = 32 - dlg;
    W.gslot = (uint32_t *)(b + o_gslot); W.grp_rec = (uint32_t *)(b + o_grec); W.grp_dslot = (uint32_t *)(b + o_gds);
    W.totals = (uint32_t *)(b + o_totals); W.ghist = (uint32_t *)(b + o_ghist);
    W.st_acc = st; W.st_gf = st + c_tiles; W.st_df = st + 2 * (size_t)c_tiles;
    W.ep_acc = kx_next_epoch(ctx); W.ep_gf = kx_next_epoch(ctx); W.ep_df = kx_next_epoch(ctx);
    W.ak = (uint32_t *)(b + o_ak); W.av = (uint32_t *)(b + o_av); W.bk = (uint32_t *)(b + o_bk); W.bv = (uint32_t *)(b + o_bv);
    W.accept_index = (uint32_t *)(b + o_acc_idx); W.group_ids = (uint32_t *)(b + o_gids);
    W.group_off = (uint32_t *)(b + o_goff); W.dev_ids = (unsigned long long *)(b + o_dids); W.dev_off = (uint32_t *)(b + o_doff);

    cudaMemcpyAsync(b + o_recs, recs, n * sizeof(kxpu_devrec), cudaMemcpyHostToDevice, ctx->stream);
    const unsigned g = (N + 255) / 256;
    uint32_t *ak = W.ak, *av = W.av, *ak2 = (uint32_t *)(b + o_ak2), *av2 = (uint32_t *)(b + o_av2);
    uint32_t *bk = W.bk, *bv = W.bv, *bk2 = (uint32_t *)(b + o_bk2), *bv2 = (uint32_t *)(b + o_bv2);
    {
        KxTimer tm(ctx, KXPU_T_CLASSIFY);
        k_reset<<<std::min<unsigned>((unsigned)((ff_bytes / 16 + 255) / 256), 8u * ctx->sm_count), 256, 0, ctx->stream>>>(
            (uint4 *)b, ff_bytes / 16, W.totals, (uint32_t)zero_words);
        k_candidates<<<g, 256, 0, ctx->stream>>>(W);
        k_accept_scan<<<c_tiles, C_THREADS, 0, ctx->stream>>>(W);
        k_groups<<<g, 256, 0, ctx->stream>>>(W);
        k_devfirst_scan<<<c_tiles, C_THREADS, 0, ctx->stream>>>(W);
        k_pairs<<<std::min<unsigned>(g, 4u * ctx->sm_count), 256, 0, ctx->stream>>>(W, passes);
        ctx->launches += 6;
        for (uint32_t p = 0; p < passes; p++) {
            SweepParams S;
            S.shift = 8 * p;
            S.epoch = kx_next_epoch(ctx);
            S.job[0] = SortJob{ak, av, ak2, av2, W.totals + 0, W.ghist + (0 * 4 + p) * 256, st + 3 * (size_t)c_tiles};
            S.job[1] = SortJob{bk, bv, bk2, bv2, W.totals + 1, W.ghist + (1 * 4 + p) * 256, st + 3 * (size_t)c_tiles + (size_t)s_tiles * 256};
            k_onesweep<<<dim3(s_tiles, 2), OS_THREADS, 0, ctx->stream>>>(S);
            ctx->launches++;
            std::swap(ak, ak2); std::swap(av, av2); std::swap(bk, bk2); std::swap(bv, bv2);
        }
        BoundsParams B;
        B.keys[0] = ak; B.count[0] = W.totals + 0; B.nord[0] = W.totals + 1; B.off[0] = W.group_off;
        B.keys[1] = bk; B.count[1] = W.totals + 1; B.nord[1] = W.totals + 2; B.off[1] = W.dev_off;
        k_bounds<<<dim3(std::min<unsigned>(g, 2u * ctx->sm_count), 2), 256, 0, ctx->stream>>>(B);
        ctx->launches++;
    }
    // results: sorted values are the CSR payloads
    uint32_t *h = ctx->h_ctl;
    cudaMemcpyAsync(h, W.totals, 16, cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    int32_t rc = KXPU_OK;
    if (e != cudaSuccess) { KX_SET_ERR(ctx, "classify failed: %s", cudaGetErrorString(e)); rc = KXPU_E_CUDA; }
    else if ((h[3] & 2u) && small_dtab) { *retry = true; return KXPU_E_CAPACITY; }
    else if (h[3] & 1u) { KX_SET_ERR(ctx, "classify: record outside the supported domain (group 0xffffffff or id file > 8 bytes)"); rc = KXPU_E_UNSUPPORTED; }
    else if (h[3] & 2u) { KX_SET_ERR(ctx, "classify: device-id table overflow"); rc = KXPU_E_CAPACITY; }
    if (rc == KXPU_OK) {
        const uint32_t na = h[0], ng = h[1], nd = h[2];
        out->n_accepted = na; out->n_groups = ng; out->n_devids = nd;
        cudaMemcpyAsync(out->accept_index, W.accept_index, n * 4, cudaMemcpyDeviceToHost, ctx->stream);
        cudaMemcpyAsync(out->group_ids, W.group_ids, (size_t)ng * 4, cudaMemcpyDeviceToHost, ctx->stream);
        cudaMemcpyAsync(out->group_off, W.group_off, ((size_t)ng + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream);
        cudaMemcpyAsync(out->group_members, av, (size_t)na * 4, cudaMemcpyDeviceToHost, ctx->stream);
        cudaMemcpyAsync(out->dev_ids, W.dev_ids, (size_t)nd * 8, cudaMemcpyDeviceToHost, ctx->stream);
        cudaMemcpyAsync(out->dev_off, W.dev_off, ((size_t)nd + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream);
        cudaMemcpyAsync(out->dev_groups, bv, (size_t)ng * 4, cudaMemcpyDeviceToHost, ctx->stream);
        e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { KX_SET_ERR(ctx, "classify D2H failed: %s", cudaGetErrorString(e)); rc = KXPU_E_CUDA; }
        if (ng == 0) out->group_off[0] = 0;
        if (nd == 0) out->dev_off[0] = 0;
    }
    return rc;
}
