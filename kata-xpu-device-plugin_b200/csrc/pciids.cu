// pciids.cu -- pci.ids parse (K1), (vendor,device) table build (K2), batched join (K3)
// and name sanitiser (K4) for sm_100a.
//
// Replaces getDeviceName / locateVendor (reference pkg/device_plugin/device_plugin.go:
// 208-275): instead of re-scanning the text once per device id, the text is streamed
// through shared memory ONCE and every (vendor,device) pair is folded into a hash table
// with "first occurrence wins" semantics; lookups are then O(1) probes.
//
// Data flow of the parse kernel (one persistent CTA per SM slot, 512 threads):
//   HBM --TMA bulk copy (cp.async.bulk, 3-stage mbarrier ring)--> 16 KiB tile in smem
//   phase a: every thread owns 2 x 16 B; SWAR newline detection -> 16-bit line-start
//            masks -> warp-level ordered compaction (shuffle scan) of line starts into a
//            per-warp list in smem; first byte peeked to find "top-level" lines.
//   phase b: one lane per line: classify (top-level / device / other), parse the 4 hex
//            digits, resolve the governing top-level line with a warp ballot
//            (segmented "last vendor" scan) + warp carry + decoupled look-back carry
//            across tiles; device lines fold into the table with atomicMin(offset).
#include "common.cuh"
#include "scan.cuh"
#include "table.cuh"

namespace kxparse {

constexpr int T = 16384;          // tile bytes
constexpr int LEAD = 16;          // bytes staged before the tile (previous byte peek)
constexpr int TRAIL = 16;         // bytes staged after the tile (line head reads)
constexpr int STAGE_BYTES = LEAD + T + TRAIL;
constexpr int NT = 512;
constexpr int NW = NT / 32;
constexpr int STAGES = 3;
constexpr int WSPAN = T / NW;     // bytes per warp = max list entries per warp

// tile_state word: [63:62] status, [61] has_top, [60] vendor valid, [59:44] vendor, [43:0] anchor
constexpr unsigned long long ST_NONE = 1ull << 62;    // aggregate only: tile holds no top-level line
constexpr unsigned long long ST_PREFIX = 2ull << 62;  // inclusive prefix
constexpr unsigned long long ST_MASK = 3ull << 62;
constexpr unsigned long long CV_HAS_TOP = 1ull << 61;
constexpr unsigned long long CV_VOK = 1ull << 60;
constexpr unsigned long long CV_ANCHOR_MASK = (1ull << 44) - 1;

struct ParseParams {
    const uint8_t *text;
    unsigned long long n;       // bytes of this shard
    unsigned long long base;    // global offset of text[0]
    uint32_t num_tiles;
    unsigned long long *tile_state;
    KxTableDev tab;
    unsigned long long carry_in;  // CV_* encoded governing line at shard start (0 = none)
};

struct ParseSmem {
    alignas(128) uint8_t stage[STAGES][STAGE_BYTES];
    alignas(16) uint16_t list[NW][WSPAN];
    alignas(8) unsigned long long full_bar[STAGES];
    unsigned long long tile_carry;
    uint32_t warp_top[NW];
    volatile uint32_t carry_seq;
    uint32_t vbid;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
// 1-D TMA bulk copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// 0x80 in every byte of x that equals '\n'
__device__ __forceinline__ uint32_t nl_flags(uint32_t x) {
    uint32_t y = x ^ 0x0a0a0a0au;
    uint32_t t = (y & 0x7f7f7f7fu) + 0x7f7f7f7fu;
    return ~(t | y | 0x7f7f7f7fu);
}
// gather the four 0x80 flags of a word into bits 0..3 (byte order)
__device__ __forceinline__ uint32_t gather4(uint32_t f) { return ((f >> 7) * 0x01020408u) >> 24; }

__device__ __forceinline__ uint32_t nl_mask16(const uint4 v) {
    return gather4(nl_flags(v.x)) | (gather4(nl_flags(v.y)) << 4) | (gather4(nl_flags(v.z)) << 8) |
           (gather4(nl_flags(v.w)) << 12);
}

__device__ __forceinline__ bool tile_uses_tma(const ParseParams &P, uint32_t t) {
    unsigned long long s = (unsigned long long)t * T;
    return s >= LEAD && s + T + TRAIL <= P.n;
}

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p) {
    return *reinterpret_cast<const volatile unsigned long long *>(p);
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long *p, unsigned long long v) {
    *reinterpret_cast<volatile unsigned long long *>(p) = v;
}

// Fold one device line into the table (first occurrence wins, see table.cuh).
__device__ __forceinline__ void table_fold(const KxTableDev &tb, uint32_t key, unsigned long long line_g,
                                           unsigned long long anchor_g) {
    uint32_t slot;
    if (key == KX_EMPTY_KEY) {
        slot = tb.cap;  // dedicated slot: 0xffffffff doubles as the empty marker
    } else {
        slot = kx_hash(key) >> tb.shift;
        uint32_t step = 0;
        for (;;) {
            uint32_t k = tb.keys[slot];
            if (k == key) break;
            if (k == KX_EMPTY_KEY) {
                uint32_t old = atomicCAS(&tb.keys[slot], KX_EMPTY_KEY, key);
                if (old == KX_EMPTY_KEY) {
                    uint32_t nk = atomicAdd(&tb.counters[KX_C_NKEYS], 1u) + 1u;
                    if (nk > tb.max_keys) tb.counters[KX_C_OVERFLOW] = 1u;
                    break;
                }
                if (old == key) break;
            }
            slot = (slot + 1) & (tb.cap - 1);
            if (++step >= tb.cap) { tb.counters[KX_C_OVERFLOW] = 1u; return; }
        }
    }
    if (line_g < tb.min_line[slot]) {
        atomicMin(&tb.min_line[slot], line_g);
        atomicMin(&tb.min_anchor[slot], anchor_g);
    }
}

__global__ void __launch_bounds__(NT, 2) parse_kernel(const ParseParams P) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    ParseSmem &S = *reinterpret_cast<ParseSmem *>(smem_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;

    if (tid == 0) {
        // virtual block id: a CTA that holds id v is resident, and so was every CTA with a
        // smaller id, which makes the look-back below deadlock free.
        S.vbid = atomicAdd(&P.tab.counters[KX_C_TICKET], 1u);
        for (int s = 0; s < STAGES; s++) mbar_init(&S.full_bar[s], 1);
        S.carry_seq = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t vbid = S.vbid, G = gridDim.x;

    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) {
            uint32_t t = vbid + (uint32_t)s * G;
            if (t < P.num_tiles && tile_uses_tma(P, t)) {
                mbar_expect_tx(&S.full_bar[s], STAGE_BYTES);
                tma_load_1d(S.stage[s], P.text + (unsigned long long)t * T - LEAD, STAGE_BYTES, &S.full_bar[s]);
            }
        }
    }

    uint32_t phase_bits = 0, it = 0;
    for (uint32_t t = vbid; t < P.num_tiles; t += G, ++it) {
        const int s = (int)(it % STAGES);
        const unsigned long long tile_start = (unsigned long long)t * T;
        uint8_t *st = S.stage[s] + LEAD;  // st[p] == text[tile_start + p]

        if (tile_uses_tma(P, t)) {
            mbar_wait(&S.full_bar[s], (phase_bits >> s) & 1u);
            phase_bits ^= 1u << s;
        } else {
            // first tile of the shard and the ragged tail: bounded loads, zero fill; the byte
            // before the shard start is a virtual '\n' so position 0 is a line start.
            for (int c = (int)tid; c < STAGE_BYTES / 16; c += NT) {
                long long g = (long long)tile_start - LEAD + 16ll * c;
                uint4 v;
                if (g >= 0 && (unsigned long long)g + 16 <= P.n) {
                    v = *reinterpret_cast<const uint4 *>(P.text + g);
                } else {
                    uint8_t tmp[16];
#pragma unroll
                    for (int b = 0; b < 16; b++) {
                        long long q = g + b;
                        tmp[b] = q < 0 ? (uint8_t)'\n' : ((unsigned long long)q < P.n ? P.text[q] : (uint8_t)0);
                    }
                    v = *reinterpret_cast<uint4 *>(tmp);
                }
                *reinterpret_cast<uint4 *>(S.stage[s] + 16 * c) = v;
            }
            __syncthreads();
        }

        // ---------------------------------------------------------------- phase a
        const unsigned long long remain = P.n - tile_start;
        const uint32_t n_rel = remain < (unsigned long long)T ? (uint32_t)remain : (uint32_t)T;
        uint32_t m[2];
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const uint32_t pos0 = (w * 64u + (uint32_t)j * 32u + lane) * 16u;
            const uint4 v = *reinterpret_cast<const uint4 *>(st + pos0);
            uint32_t nl = nl_mask16(v);
            uint32_t prev_nl = st[(int)pos0 - 1] == (uint8_t)'\n';
            uint32_t ls = ((nl << 1) | prev_nl) & 0xffffu;  // line start <=> previous byte is '\n'
            if (pos0 + 16u > n_rel) ls &= pos0 >= n_rel ? 0u : ((1u << (n_rel - pos0)) - 1u);
            m[j] = ls;
        }
        const uint32_t cnt = (uint32_t)__popc(m[0]) | ((uint32_t)__popc(m[1]) << 16);
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= (uint32_t)d) incl += y;
        }
        const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
        const uint32_t tot0 = tot & 0xffffu, L = tot0 + (tot >> 16);
        const uint32_t excl = incl - cnt;
        uint32_t last_top = 0;  // position+1 of the last top-level line start seen by this lane
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const uint32_t pos0 = (w * 64u + (uint32_t)j * 32u + lane) * 16u;
            uint32_t idx = j == 0 ? (excl & 0xffffu) : tot0 + (excl >> 16);
            uint32_t mm = m[j];
            while (mm) {
                uint32_t b = (uint32_t)__ffs((int)mm) - 1u;
                mm &= mm - 1u;
                uint32_t p = pos0 + b;
                uint8_t c0 = st[p];
                // a line that starts with neither '#' nor '\t' ends the current vendor block
                // (device_plugin.go:229-236) and is the only kind locateVendor can match (:265)
                uint32_t top = (c0 != (uint8_t)'#') & (c0 != (uint8_t)'\t');
                S.list[w][idx++] = (uint16_t)(p | (top << 15));
                if (top) last_top = p + 1u;
            }
        }
        const uint32_t wlt = __reduce_max_sync(0xffffffffu, last_top);
        if (lane == 0) {
            uint32_t Pk = 0;
            if (wlt) {
                uint32_t p = wlt - 1u, v;
                uint32_t h = (uint32_t)st[p] | ((uint32_t)st[p + 1] << 8) | ((uint32_t)st[p + 2] << 16) |
                             ((uint32_t)st[p + 3] << 24);
                bool ok = kx_hex4(h, v);
                Pk = 0x80000000u | (ok ? 0x40000000u : 0u) | (v << 14) | p;
            }
            S.warp_top[w] = Pk;
        }
        __syncthreads();  // #A: lists and warp_top complete

        // ---------------------------------------------------------------- phase b
        // governing top-level line carried into this warp's span: last one of the previous
        // warps of this tile, else (0) the tile's carry-in.
        uint32_t cP;
        {
            uint32_t x = lane < NW ? S.warp_top[lane] : 0u;
            uint32_t allm = __ballot_sync(0xffffffffu, x != 0u);
            uint32_t prevm = allm & ((1u << w) - 1u);
            uint32_t src = prevm ? 31u - (uint32_t)__clz((int)prevm) : 0u;
            uint32_t g = __shfl_sync(0xffffffffu, x, src);
            cP = prevm ? g : 0u;
            if (w == 0) {
                uint32_t asrc = allm ? 31u - (uint32_t)__clz((int)allm) : 0u;
                uint32_t agg = __shfl_sync(0xffffffffu, x, asrc);
                if (lane == 0) {
                    unsigned long long own = 0;
                    if (allm) {
                        own = CV_HAS_TOP | ((agg & 0x40000000u) ? CV_VOK : 0ull) |
                              ((unsigned long long)((agg >> 14) & 0xffffu) << 44) |
                              ((P.base + tile_start + (agg & 0x3fffu)) & CV_ANCHOR_MASK);
                        st_volatile_u64(&P.tile_state[t], ST_PREFIX | own);
                    } else if (t > 0) {
                        st_volatile_u64(&P.tile_state[t], ST_NONE);
                    }
                    // decoupled look-back for the carry-in of this tile
                    unsigned long long carry = P.carry_in;
                    if (t > 0) {
                        uint32_t j = t - 1;
                        for (;;) {
                            unsigned long long sv = ld_volatile_u64(&P.tile_state[j]);
                            unsigned long long stt = sv & ST_MASK;
                            if (stt == ST_PREFIX) { carry = sv & ~ST_MASK; break; }
                            if (stt == ST_NONE) { j--; continue; }  // tile 0 always publishes a prefix
                            __nanosleep(20);
                        }
                    }
                    if (!allm) st_volatile_u64(&P.tile_state[t], ST_PREFIX | carry);
                    S.tile_carry = carry;
                    __threadfence_block();
                    S.carry_seq = it + 1u;
                }
            }
        }

        for (uint32_t r = 0; r < L; r += 32u) {
            const uint32_t i = r + lane;
            const bool active = i < L;
            const uint32_t e = active ? (uint32_t)S.list[w][i] : 0u;
            const uint32_t p = e & 0x3fffu;
            const bool istop = active && (e >> 15);
            // 8 bytes at st[p] (unaligned): three aligned words + funnel shifts
            const uint32_t a = (uint32_t)LEAD + p;
            const uint32_t *wp = reinterpret_cast<const uint32_t *>(S.stage[s] + (a & ~3u));
            const uint32_t sh = (a & 3u) * 8u;
            const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2];
            const uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
            const uint32_t h = istop ? lo : __funnelshift_r(lo, hi, 8);
            uint32_t val;
            const bool ok = kx_hex4(h, val);
            const bool isdev = active && !istop && (lo & 0xffu) == (uint32_t)'\t' && ok;
            const uint32_t myP = istop ? (0x80000000u | (ok ? 0x40000000u : 0u) | (val << 14) | p) : 0u;
            const uint32_t topm = __ballot_sync(0xffffffffu, istop);
            const uint32_t prev = topm & lt_mask;
            const uint32_t src = prev ? 31u - (uint32_t)__clz((int)prev) : 0u;
            const uint32_t g = __shfl_sync(0xffffffffu, myP, src);
            const uint32_t gov = prev ? g : cP;
            if (topm) cP = __shfl_sync(0xffffffffu, myP, 31u - (uint32_t)__clz((int)topm));

            const unsigned long long line_g = P.base + tile_start + p;
            if (istop && ok) {
                // candidate vendor anchor: only the first line with this prefix counts (:265)
                if (line_g < P.tab.vendor_first[val]) atomicMin(&P.tab.vendor_first[val], line_g);
            }
            const bool need_carry = isdev && !(gov >> 31);
            if (__any_sync(0xffffffffu, need_carry)) {
                while (S.carry_seq != it + 1u) { /* warp 0 lane 0 is resolving the look-back */ }
            }
            if (isdev) {
                bool vok;
                uint32_t V;
                unsigned long long anchor_g;
                if (gov >> 31) {
                    vok = (gov >> 30) & 1u;
                    V = (gov >> 14) & 0xffffu;
                    anchor_g = P.base + tile_start + (gov & 0x3fffu);
                } else {
                    unsigned long long cv = S.tile_carry;
                    vok = (cv & CV_HAS_TOP) && (cv & CV_VOK);
                    V = (uint32_t)(cv >> 44) & 0xffffu;
                    anchor_g = cv & CV_ANCHOR_MASK;
                }
                if (vok) table_fold(P.tab, (V << 16) | val, line_g, anchor_g);
            }
        }
        // A line of >= 64 KiB ends the reference's scan (bufio.ErrTooLong).  Such a line
        // leaves at least 62 whole warp spans without a line start, so a line-free span
        // that lies inside the text raises a hint; the exact cut-off is then computed by
        // trunc_kernel (never needed for real pci.ids data).
        if (L == 0 && (w + 1u) * (uint32_t)WSPAN <= n_rel && lane == 0)
            atomicOr(&P.tab.counters[KX_C_LONGLINE_HINT], 1u);
        __syncthreads();  // #B: everybody is done with stage s
        if (tid == 0) {
            uint32_t nt = t + (uint32_t)STAGES * G;
            if (nt < P.num_tiles && tile_uses_tma(P, nt)) {
                mbar_expect_tx(&S.full_bar[s], STAGE_BYTES);
                tma_load_1d(S.stage[s], P.text + (unsigned long long)nt * T - LEAD, STAGE_BYTES, &S.full_bar[s]);
            }
        }
    }
}


// ------------------------------------------------------------------------------
// bufio.ErrTooLong cut-off (slow path, only when the parse kernel raised the hint).
// trunc = global offset of the first line whose content is >= 65536 bytes.
// One CTA; every thread scans a contiguous byte range for newlines and reports the
// first/last newline and the longest gap inside; thread 0 stitches the ranges.
// ------------------------------------------------------------------------------
constexpr unsigned long long MAX_TOKEN = 65536ull;

__global__ void __launch_bounds__(1024) trunc_kernel(const uint8_t *__restrict__ text, unsigned long long n,
                                                      unsigned long long base, unsigned long long *trunc_out) {
    __shared__ unsigned long long s_first[1024], s_last[1024], s_bad[1024];
    const unsigned long long per = (n + 1023ull) / 1024ull;
    const unsigned long long lo = per * threadIdx.x, hi = lo + per < n ? lo + per : n;
    unsigned long long first = KX_NO_OFF, last = KX_NO_OFF, bad = KX_NO_OFF;
    for (unsigned long long i = lo; i < hi; i++) {
        if (text[i] == (uint8_t)'\n') {
            if (first == KX_NO_OFF) first = i;
            else if (i - last - 1 >= MAX_TOKEN && bad == KX_NO_OFF) bad = last + 1;  // line (last, i)
            last = i;
        }
    }
    s_first[threadIdx.x] = first; s_last[threadIdx.x] = last; s_bad[threadIdx.x] = bad;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long prev_nl = KX_NO_OFF;  // offset of the last newline so far (none: line starts at 0)
        unsigned long long res = KX_NO_OFF;
        for (int k = 0; k < 1024 && res == KX_NO_OFF; k++) {
            if (s_first[k] != KX_NO_OFF) {
                unsigned long long start = prev_nl == KX_NO_OFF ? 0 : prev_nl + 1;
                if (s_first[k] - start >= MAX_TOKEN) { res = start; break; }
                if (s_bad[k] != KX_NO_OFF) { res = s_bad[k]; break; }
                prev_nl = s_last[k];
            }
        }
        if (res == KX_NO_OFF) {
            unsigned long long start = prev_nl == KX_NO_OFF ? 0 : prev_nl + 1;
            if (n - start >= MAX_TOKEN) res = start;  // unterminated final line
        }
        *trunc_out = res == KX_NO_OFF ? KX_NO_OFF : base + res;
    }
}

// ------------------------------------------------------------------------------
// K4 name sanitiser (device_plugin.go:241-251), warp-cooperative.
// ------------------------------------------------------------------------------
__device__ __forceinline__ bool is_re_space(uint32_t c) {  // RE2 \s
    return c == 0x20u || c == 0x09u || c == 0x0au || c == 0x0cu || c == 0x0du;
}
// length of a unicode.IsSpace rune starting at s[0] (0 = not a space); len = bytes available
__device__ __forceinline__ uint32_t uspace_len(const uint8_t *s, uint32_t len) {
    if (len == 0) return 0;
    uint32_t c = s[0];
    if (c == 0x20u || (c >= 0x09u && c <= 0x0du)) return 1;
    if (len >= 2 && c == 0xC2u && (s[1] == 0x85u || s[1] == 0xA0u)) return 2;
    if (len >= 3) {
        uint32_t d = s[1], e = s[2];
        if (c == 0xE1u && d == 0x9Au && e == 0x80u) return 3;
        if (c == 0xE2u && d == 0x80u && ((e >= 0x80u && e <= 0x8Au) || e == 0xA8u || e == 0xA9u || e == 0xAFu)) return 3;
        if (c == 0xE2u && d == 0x81u && e == 0x9Fu) return 3;
        if (c == 0xE3u && d == 0x80u && e == 0x80u) return 3;
    }
    return 0;
}
// strings.TrimSpace on buf[0..len): returns [start,end)
__device__ __forceinline__ void trim_space(const uint8_t *buf, uint32_t len, uint32_t &start, uint32_t &end) {
    uint32_t a = 0, b = len, k;
    while ((k = uspace_len(buf + a, b - a)) != 0) a += k;
    for (;;) {
        if (b > a && uspace_len(buf + b - 1, 1) == 1) { b -= 1; continue; }
        if (b - a >= 2 && uspace_len(buf + b - 2, 2) == 2) { b -= 2; continue; }
        if (b - a >= 3 && uspace_len(buf + b - 3, 3) == 3) { b -= 3; continue; }
        break;
    }
    start = a; end = b;
}
// output byte for position i of the trimmed range (0 = deleted)
__device__ __forceinline__ uint32_t sanitise_byte(const uint8_t *buf, uint32_t i, uint32_t start, uint32_t end) {
    uint32_t c = buf[i];
    if (is_re_space(c)) return (i > start && is_re_space(buf[i - 1])) ? 0u : (uint32_t)'_';
    if (c >= 'a' && c <= 'z') return c - 32u;
    if (c == '/' || c == '.') return (uint32_t)'_';
    if ((c >= 'A' && c <= 'Z') || (c >= '0' && c <= '9') || c == '_') return c;
    if (i + 1 < end) {
        if (c == 0xC4u && buf[i + 1] == 0xB1u) return (uint32_t)'I';  // U+0131 upper-cases to ASCII I
        if (c == 0xC5u && buf[i + 1] == 0xBFu) return (uint32_t)'S';  // U+017F upper-cases to ASCII S
    }
    return 0u;
}

constexpr int NAME_BUF = 1024;  // fast path: rest-of-line fits the per-warp staging buffer
constexpr int FIN_WARPS = 8;

struct FinalizeParams {
    const uint8_t *text;  // shard text (local)
    unsigned long long n, base;
    KxTableDev tab;
    int32_t *row_of_slot;
    uint32_t *row_key;
    unsigned long long *row_line;
    unsigned long long *row_anchor;
    uint32_t *row_name_off;
    uint32_t *row_name_len;
    uint32_t *sel;  // [cap+1] valid slots (stage 1 -> stage 2)
    uint8_t *blob;
    uint32_t blob_cap;
    int check_valid;  // 1: apply vendor_first / trunc validity (single shard); 0: emit every local row (sharded)
};

// Stage 1, one thread per table slot: validity; valid slots are compacted into F.sel.
__global__ void __launch_bounds__(256) finalize_select_kernel(const FinalizeParams F) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (F.tab.counters[KX_C_LONGLINE_HINT] != 0u && F.tab.counters[KX_C_NEED_TRUNC] == 0u && F.check_valid) {
        // the exact ErrTooLong cut-off has not been computed yet: ask the host to run
        // trunc_kernel and call finalize again.
        if (slot == 0) F.tab.counters[KX_C_NEED_TRUNC] = 2u;
        return;
    }
    bool valid = false;
    if (slot <= F.tab.cap) {
        const unsigned long long line = F.tab.min_line[slot];
        const uint32_t key = slot == F.tab.cap ? KX_EMPTY_KEY : F.tab.keys[slot];
        valid = line != KX_NO_OFF && !(slot < F.tab.cap && key == KX_EMPTY_KEY);
        if (valid && F.check_valid) valid = F.tab.min_anchor[slot] == F.tab.vendor_first[key >> 16] && line < *F.tab.trunc;
        if (!valid) F.row_of_slot[slot] = -1;
    }
    const uint32_t vm = __ballot_sync(0xffffffffu, valid);
    if (vm) {
        uint32_t base = 0;
        if ((threadIdx.x & 31u) == 0) base = atomicAdd(&F.tab.counters[KX_C_NSEL], (uint32_t)__popc(vm));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (valid) F.sel[base + (uint32_t)__popc(vm & ((1u << (threadIdx.x & 31u)) - 1u))] = slot;
    }
}

// Stage 2, one warp per selected slot (row handle = index in F.sel): sanitised name into the
// blob.  Blob space is claimed once per CTA (8 names) to keep the cursor atomic cheap.
__global__ void __launch_bounds__(FIN_WARPS * 32) finalize_kernel(const FinalizeParams F) {
    __shared__ uint8_t s_buf[FIN_WARPS][NAME_BUF + 32];
    __shared__ uint32_t s_len[FIN_WARPS];
    __shared__ uint32_t s_base;
    const uint32_t lane = threadIdx.x & 31u, wl = threadIdx.x >> 5;
    const uint32_t nsel = F.tab.counters[KX_C_NSEL];
    const uint32_t si = blockIdx.x * FIN_WARPS + wl;
    if (blockIdx.x * FIN_WARPS >= nsel) return;  // whole CTA idle
    const bool active = si < nsel;
    uint32_t slot = 0, key = 0, len = 0, start = 0, end = 0, out_len = 0;
    unsigned long long line = 0, anchor = 0, rs = 0;
    bool fast = false;
    uint8_t *buf = s_buf[wl];
    if (active) {
        slot = F.sel[si];
        line = F.tab.min_line[slot];
        key = slot == F.tab.cap ? KX_EMPTY_KEY : F.tab.keys[slot];
        anchor = F.tab.min_anchor[slot];
        rs = line - F.base + 5ull;  // rest of the line after "\t" + 4 hex digits
        bool found = false;
        for (uint32_t o = 0; o < (uint32_t)NAME_BUF + 32u && !found; o += 32u) {
            unsigned long long pos = rs + o + lane;
            uint32_t c = pos < F.n ? F.text[pos] : 0x0au;  // EOF terminates the last line
            buf[o + lane] = (uint8_t)c;
            uint32_t nlm = __ballot_sync(0xffffffffu, c == 0x0au);
            if (nlm) { len = o + (uint32_t)__ffs((int)nlm) - 1u; found = true; }
        }
        __syncwarp();
        fast = found && len <= (uint32_t)NAME_BUF;
        if (fast) {
            if (len > 0 && buf[len - 1] == 0x0du) len--;  // bufio.ScanLines drops one trailing CR
            if (lane == 0) trim_space(buf, len, start, end);
            start = __shfl_sync(0xffffffffu, start, 0);
            end = __shfl_sync(0xffffffffu, end, 0);
            for (uint32_t o = start; o < end; o += 32u) {
                uint32_t i = o + lane;
                uint32_t ch = i < end ? sanitise_byte(buf, i, start, end) : 0u;
                out_len += (uint32_t)__popc(__ballot_sync(0xffffffffu, ch != 0u));
            }
        } else {
            // slow path: a name longer than the staging buffer (never in pci.ids): lane 0, serial,
            // straight from global memory.
            if (lane == 0) {
                const uint8_t *g = F.text + rs;
                unsigned long long avail = F.n - rs, l = 0;
                while (l < avail && g[l] != 0x0au) l++;
                len = (uint32_t)l;
                if (len > 0 && g[len - 1] == 0x0du) len--;
                trim_space(g, len, start, end);
                for (uint32_t i = start; i < end; i++) out_len += sanitise_byte(g, i, start, end) != 0u;
            }
            out_len = __shfl_sync(0xffffffffu, out_len, 0);
            start = __shfl_sync(0xffffffffu, start, 0);
            end = __shfl_sync(0xffffffffu, end, 0);
        }
    }
    if (lane == 0) s_len[wl] = out_len;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tot = 0;
        for (int k = 0; k < FIN_WARPS; k++) tot += s_len[k];
        uint32_t base = tot ? atomicAdd(&F.tab.counters[KX_C_BLOB_CURSOR], tot) : 0u;
        if (base + tot > F.blob_cap) { F.tab.counters[KX_C_BLOB_OVERFLOW] = 1u; base = 0xFFFFFFFFu; }
        s_base = base;
    }
    __syncthreads();
    if (!active) return;
    uint32_t out_off = s_base;
    const bool room = out_off != 0xFFFFFFFFu;
    for (uint32_t k = 0; k < wl; k++) out_off += s_len[k];
    if (room) {
        if (fast) {
            uint32_t wr = out_off;
            for (uint32_t o = start; o < end; o += 32u) {
                uint32_t i = o + lane;
                uint32_t ch = i < end ? sanitise_byte(buf, i, start, end) : 0u;
                uint32_t bm = __ballot_sync(0xffffffffu, ch != 0u);
                if (ch) F.blob[wr + (uint32_t)__popc(bm & ((1u << lane) - 1u))] = (uint8_t)ch;
                wr += (uint32_t)__popc(bm);
            }
        } else if (lane == 0) {
            const uint8_t *g = F.text + rs;
            uint32_t wr = out_off;
            for (uint32_t i = start; i < end; i++) {
                uint32_t ch = sanitise_byte(g, i, start, end);
                if (ch) F.blob[wr++] = (uint8_t)ch;
            }
        }
    }
    if (lane == 0) {
        F.row_of_slot[slot] = (int32_t)si;
        F.row_key[si] = key;
        F.row_line[si] = line;
        F.row_anchor[si] = anchor;
        F.row_name_off[si] = room ? out_off : 0u;
        F.row_name_len[si] = room ? out_len : 0u;
    }
}

// ------------------------------------------------------------------------------
// K3 batched join: one thread per key, probe the table (L2 resident), return row handle.
// ------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) lookup_kernel(const uint32_t *__restrict__ keys, size_t n,
                                                      const uint32_t *__restrict__ tkeys,
                                                      const int32_t *__restrict__ row_of_slot, uint32_t cap,
                                                      uint32_t shift, int32_t *__restrict__ rows_out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const uint32_t key = keys[i];
        int32_t row = -1;
        if (key == KX_EMPTY_KEY) {
            row = row_of_slot[cap];
        } else {
            uint32_t slot = kx_hash(key) >> shift;
            for (uint32_t step = 0; step < cap; step++) {
                uint32_t k = __ldg(&tkeys[slot]);
                if (k == key) { row = __ldg(&row_of_slot[slot]); break; }
                if (k == KX_EMPTY_KEY) break;
                slot = (slot + 1) & (cap - 1);
            }
        }
        rows_out[i] = row;
    }
}

// name gather: lengths, then copy
__global__ void __launch_bounds__(256) name_len_kernel(const int32_t *__restrict__ rows, size_t n,
                                                        const uint32_t *__restrict__ row_name_len, uint32_t n_rows,
                                                        uint32_t *__restrict__ lens) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        int32_t r = rows[i];
        lens[i] = (r >= 0 && (uint32_t)r < n_rows) ? row_name_len[r] : 0u;
    }
}
// 8 lanes per name
__global__ void __launch_bounds__(256) name_copy_kernel(const int32_t *__restrict__ rows, size_t n,
                                                         const uint32_t *__restrict__ row_name_off,
                                                         const uint32_t *__restrict__ row_name_len, uint32_t n_rows,
                                                         const uint8_t *__restrict__ blob,
                                                         const uint32_t *__restrict__ offsets, uint8_t *__restrict__ out,
                                                         size_t cap) {
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    uint32_t sub = threadIdx.x & 7u;
    if (i >= n) return;
    int32_t r = rows[i];
    if (r < 0 || (uint32_t)r >= n_rows) return;
    uint32_t len = row_name_len[r], src = row_name_off[r], dst = offsets[i];
    if ((size_t)dst + len > cap) return;
    for (uint32_t k = sub; k < len; k += 8u) out[dst + k] = blob[src + k];
}

}  // namespace kxparse
