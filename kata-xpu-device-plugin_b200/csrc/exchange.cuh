// exchange.cuh -- device-side pieces of the sharded load's exchange that ride on kernels of the local
// pipeline (comm.cu owns the protocol): the phase-A push, run by the LAST CTA of the resolve kernel,
// and the flag wait that a consumer kernel can run in its prologue.
#pragma once
#include "common.cuh"
#include "table.cuh"

namespace kxx {

constexpr uint32_t XS_GROW = 1u;           // a rank's table overflowed / is too full
constexpr uint32_t XS_NEED_TRUNC = 2u;     // a rank saw a possible >= 64 KiB line: run again with exact cut-offs
constexpr uint32_t XS_SLAB_OVERFLOW = 4u;  // a rank's winners outgrow the slab
constexpr uint32_t XS_GROW_BLOB = 8u;      // a rank's name blob is too small
constexpr uint32_t XS_FULL = 16u;          // a rank's table ran completely full (its key count is unknown): grow faster

// phase-A block of ONE source rank inside a region: u64 [65536] vendor_first | cut-off | status | pad.
// Every rank owns one such block per buffer in every region and overwrites it completely each epoch
// (plain 16-byte stores, nothing to clear, no remote atomics); consumers take the minimum over the
// blocks of all ranks on the fly.
constexpr int A_TRUNC = 65536;
constexpr int A_STATUS = 65537;  // XS_* bits of the source rank (plain word)
constexpr int A_WORDS = 65536 + 8;

// the R phase-A blocks of one buffer as the consumers see them
struct MinView {
    const unsigned long long *a;  // block of rank 0 (single text: the table's own vendor_first, n = 1)
    size_t stride;                // words between the blocks of consecutive ranks
    int n;
    const unsigned long long *trunc1;  // n == 1 only: the table's own cut-off word
};
__device__ __forceinline__ unsigned long long min_view_first(const MinView &V, uint32_t vendor) {
    unsigned long long m = V.a[vendor];
    for (int r = 1; r < V.n; r++) {
        const unsigned long long x = V.a[(size_t)r * V.stride + vendor];
        m = x < m ? x : m;
    }
    return m;
}
__device__ __forceinline__ unsigned long long min_view_trunc(const MinView &V) {
    if (V.n == 1 && V.trunc1) return *V.trunc1;
    unsigned long long m = KX_NO_OFF;
    for (int r = 0; r < V.n; r++) {
        const unsigned long long x = V.a[(size_t)r * V.stride + A_TRUNC];
        m = x < m ? x : m;
    }
    return m;
}
__device__ __forceinline__ uint32_t min_view_status(const MinView &V) {
    uint32_t st = 0;
    for (int r = 0; r < V.n; r++) st |= (uint32_t)V.a[(size_t)r * V.stride + A_STATUS];
    return st;
}

struct SlabHeader {  // 64 bytes
    uint32_t n_rows, blob_bytes, status, nkeys;
    uint32_t pad[12];
};
struct SlabRow {  // 32 bytes: one winner (vendor,device) row of the shard
    uint32_t key, name_len;
    unsigned long long line, anchor;
    uint32_t name_off, pad;
};

struct Targets {  // where a push goes: every rank's region (peer memory) or this rank's staging region (NCCL)
    uint8_t *region[KX_MAX_RANKS];
    int n;
};

// Phase B is a PULL: the finalize writes the winners into this rank's own slab; its last CTA adds the header and
// raises "slab ready" on every peer; a peer's merge kernel reads the slab over NVLink.  (A push kernel in
// between cost 11-18 us per load for the launch, one system fence per CTA and two in the last one.)
struct SlabTail {
    int on;
    uint32_t *done;          // last-CTA counter
    SlabHeader *header;      // of my own slab (local memory)
    Targets tg;              // regions that hold the flags (peer transport), n = 0: no flags (NCCL transport)
    size_t o_flag;
    uint32_t epoch;
    uint32_t rows_cap, blob_cap;
};

struct XaParams {
    Targets tg;
    size_t o_a;     // my phase-A block inside a region
    size_t o_flag;  // my phase-A flag inside a region (peer transport)
    int raise_flags;
    uint32_t epoch;
    const unsigned long long *vendor_first, *trunc;
    const uint32_t *counters;
    uint32_t max_keys;
    int have_trunc;
};

// Phase A: this shard's vendor_first (dense, 512 KB) goes into my block of every rank's region with
// plain 16-byte stores over NVLink.  vendor_first is final before the last resolve kernel starts, so
// ALL its threads share the copy (xa_push_slice); cut-off, status and the flags wait for the CTA that
// finishes last (xa_finish, thread 0 of that CTA, after every CTA's fence + counter).
__device__ __forceinline__ void xa_push_slice(const XaParams &P, uint32_t gtid, uint32_t gthreads) {
    const uint4 *src = reinterpret_cast<const uint4 *>(P.vendor_first);
    for (uint32_t i = gtid; i < 65536u / 2u; i += gthreads) {
        const uint4 x = src[i];
        for (int q = 0; q < P.tg.n; q++) reinterpret_cast<uint4 *>(P.tg.region[q] + P.o_a)[i] = x;
    }
}
__device__ __forceinline__ void xa_finish(const XaParams &P) {
    const unsigned long long t = *reinterpret_cast<const volatile unsigned long long *>(P.trunc);
    const volatile uint32_t *c = P.counters;
    uint32_t st = 0;
    if (c[KX_C_OVERFLOW]) st |= XS_GROW | XS_FULL;
    if (c[KX_C_NKEYS] > P.max_keys) st |= XS_GROW;
    if (c[KX_C_LONGLINE_HINT] && !P.have_trunc) st |= XS_NEED_TRUNC;
    for (int q = 0; q < P.tg.n; q++) {
        unsigned long long *a = reinterpret_cast<unsigned long long *>(P.tg.region[q] + P.o_a);
        a[A_TRUNC] = t;
        a[A_STATUS] = st;
    }
    kx_fence_sys();
    if (P.raise_flags)
        for (int q = 0; q < P.tg.n; q++) *reinterpret_cast<volatile uint32_t *>(P.tg.region[q] + P.o_flag) = P.epoch;
}

// Flags of one phase: lane q of the calling warp waits for rank q's flag of this epoch (~4 s time-out).
struct WaitSpec {
    const uint32_t *flags;  // nullptr: nothing to wait for
    int nranks;
    uint32_t epoch;
    uint32_t *timeout_flag;
};
__device__ __forceinline__ void wait_flags_lane(const WaitSpec &W, int q) {
    if (q < W.nranks) {
        const long long t0 = clock64();
        while (*reinterpret_cast<const volatile uint32_t *>(&W.flags[q]) != W.epoch) {
            __nanosleep(64);
            if (clock64() - t0 > 8000000000ll) { *W.timeout_flag = 1u; break; }  // a peer died or the ranks lost step
        }
    }
    kx_fence_sys();
}
// prologue of a consumer kernel (every CTA): the first warp waits, the barrier releases the others
__device__ __forceinline__ void wait_flags_cta(const WaitSpec &W) {
    if (W.flags == nullptr) return;
    if (threadIdx.x < 32) wait_flags_lane(W, (int)threadIdx.x);
    __syncthreads();
}

}  // namespace kxx
