// pciids3.cu -- parse kernel v3: super-chunk streaming parse of pci.ids text.
//
// Lessons of v1/v2 (profiles/r01_parse_v1_ncu_summary.txt, r01_parse_v2b_ncu_summary.txt):
// the per-chunk work of v2 runs at ~2.6 TB/s, but 55 % of its instructions were warps polling
// a per-chunk status word in global memory: 70 % of all device lines are governed by a vendor
// line in an EARLIER 2 KiB chunk, and inside the big vendor blocks (8086: 330 KB) a chunk can
// only be resolved when ~100 other warps on other SMs have published.  v3 keeps the per-chunk
// machinery of v2 and changes who talks to whom:
//   * a CTA (8 warps) owns a SUPER-CHUNK of 32 consecutive 2 KiB chunks (64 KiB); warp w parses
//     chunks w, w+8, w+16, w+24 of it from its private 2-stage TMA ring (cp.async.bulk + mbarrier);
//   * the "current vendor" carry between chunks of a super-chunk goes through a 32-entry status
//     array in SHARED memory (published after the cheap first pass, consumed at the end of the
//     chunk: practically never waited for);
//   * lines governed by a vendor line of an earlier super-chunk are PARKED in a CTA buffer
//     (device id + position, 4 B each) and folded one super-chunk iteration later, after one
//     warp resolved the carry with a decoupled look-back over per-SUPER-CHUNK status words in
//     global memory (32x fewer words, a whole iteration of slack);
//   * one split barrier (mbarrier arrive / wait) per 64 KiB.
#pragma once
#include "common.cuh"
#include "pciids2.cu"  // nl_mask16, hex4_swar, table_fold, TMA/mbarrier helpers, CV_*/ST_* encodings
#include "table.cuh"

namespace kxparse3 {

using namespace kxparse2;  // helpers and the CV_* / ST_* status-word encoding

constexpr int STAGES3 = 2;
constexpr int SCC = 32;                       // chunks per super-chunk
constexpr int SCB = SCC * CW;                 // 65536 bytes
constexpr int CPW = SCC / WARPS;              // chunks per warp per super-chunk
constexpr int PENDCAP = 1536;                 // parked lines per super-chunk

// shared-memory chunk status (u32): [31] published, [30] has top-level line, [29] vendor ok,
// [27:12] vendor, [11:0] position of the chunk's last top-level line inside the chunk
#define LS_PUB 0x80000000u
#define LS_TOP 0x40000000u
#define LS_VOK 0x20000000u

struct WarpSmem3 {
    alignas(16) uint8_t stage[STAGES3][STG_BYTES];
    alignas(16) uint16_t list[LCAP];
    alignas(8) unsigned long long bar[STAGES3];
};

struct ChunkCarry {            // governing line of the head lines of one chunk
    unsigned long long anchor; // global offset of the governing top-level line
    uint32_t key_hi;           // vendor << 16
    uint32_t valid;            // a governing line exists and its first four bytes are lowercase hex
};

struct CtaSmem3 {
    WarpSmem3 w[WARPS];
    // iteration k parks into pend[k % 3]; warp 0 resolves its carry at the END of iteration k + 1
    // (-> ccarry[k & 1], fold_cnt[k & 1]); all warps fold it in the MIDDLE of iteration k + 2.
    alignas(16) uint32_t pend[3][PENDCAP];        // (device << 16) | newline position in the super-chunk
    alignas(16) ChunkCarry ccarry[2][SCC];
    uint32_t cstate[3][SCC];                      // see LS_*
    uint32_t pend_cnt[3];
    uint32_t agg_none[3];                         // super-chunk published ST_NONE (no top-level line)
    uint32_t sc_q[3];                             // super-chunk tickets: [k % 3] = iteration k, [(k+1) % 3] = k + 1
    uint32_t pub_cnt[2];
    uint32_t fold_cnt[2];
    alignas(8) unsigned long long iter_bar;       // mbarrier: one arrival per warp per iteration
};

struct Params3 {
    const uint8_t *text;
    unsigned long long n, base;
    uint32_t num_chunks, num_sc;
    unsigned long long *sc_state;  // [num_sc], zero initialised
    KxTableDev tab;
    unsigned long long carry_in;
};

// decoupled look-back over super-chunk status words (one warp)
__device__ __forceinline__ unsigned long long lookback_sc(const Params3 &P, uint32_t sc, uint32_t lane) {
    long long top = (long long)sc - 1;
    for (;;) {
        const long long idx = top - (long long)lane;
        const unsigned long long s = idx >= 0 ? ld_volatile_u64(&P.sc_state[idx]) : (ST_PREFIX | P.carry_in);
        const unsigned long long st = s & ST_MASK;
        const uint32_t pm = __ballot_sync(0xffffffffu, st == ST_PREFIX);
        const uint32_t zm = __ballot_sync(0xffffffffu, st == 0ull);
        if (pm) {
            const uint32_t f = (uint32_t)__ffs((int)pm) - 1u;
            if ((zm & ((1u << f) - 1u)) == 0u) return __shfl_sync(0xffffffffu, s, (int)f) & ~ST_MASK;
        } else if (zm == 0u) {
            top -= 32;
            continue;
        }
        __nanosleep(500);
    }
}

__global__ void __launch_bounds__(NT, 4) parse_kernel_v3(const Params3 P) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    CtaSmem3 &C = *reinterpret_cast<CtaSmem3 *>(smem_raw);
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    const uint32_t wl = (uint32_t)WARPS - 1u - w;
    WarpSmem3 &S = C.w[w];
    const uint32_t lt_mask = (1u << lane) - 1u;

    if (threadIdx.x == 0) {
        // Super-chunks are handed out by ticket, two in flight per CTA (current + prefetched
        // next).  A CTA that holds ticket t is resident and every smaller ticket was taken by a
        // resident CTA before: the smallest unresolved super-chunk always belongs to a running
        // CTA, so the look-back below cannot deadlock, and fast CTAs simply take more tickets.
        C.sc_q[0] = atomicAdd(&P.tab.counters[KX_C_TICKET], 1u);
        C.sc_q[1] = atomicAdd(&P.tab.counters[KX_C_TICKET], 1u);
        C.pend_cnt[0] = C.pend_cnt[1] = C.pend_cnt[2] = 0;
        C.pub_cnt[0] = C.pub_cnt[1] = 0;
        mbar_init(&C.iter_bar, WARPS);
        C.fold_cnt[0] = C.fold_cnt[1] = 0;
    }
    if (lane == 0) {
        for (int s = 0; s < STAGES3; s++) mbar_init(&S.bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    uint32_t k7f = 0x7f7f7f7fu, k0a = 0x0a0a0a0au, k80 = 0x80808080u;
    asm volatile("" : "+r"(k7f), "+r"(k0a), "+r"(k80));

    // chunks [0, tma_limit) can be staged with one bulk copy of STG_BYTES (no read past the text)
    const uint32_t tma_limit = P.n >= (unsigned long long)STG_BYTES ? (uint32_t)((P.n - STG_BYTES) / CW) + 1u : 0u;
    auto chunk_tma_ok = [&](uint32_t g) { return g < tma_limit; };
    const unsigned long long pol = l2_evict_first_policy();
    // lane 0: start the TMA copy of chunk c of super-chunk scn into stage s (if it exists)
    auto issue = [&](uint32_t scn, uint32_t c, int s) {
        if (scn >= P.num_sc) return;
        const uint32_t g = scn * SCC + c;
        if (g < P.num_chunks && chunk_tma_ok(g)) {
            mbar_expect_tx(&S.bar[s], STG_BYTES);
            tma_load_1d_stream(S.stage[s], P.text + (unsigned long long)g * CW, STG_BYTES, &S.bar[s], pol);
        }
    };
    if (lane == 0) {
        issue(C.sc_q[0], wl, 0);
        issue(C.sc_q[0], wl + WARPS, 1);
    }

    // Parked lines are resolved with two iterations of slack, never on the critical path:
    //   iteration m     : lines governed by an earlier super-chunk are parked in pend[m % 3];
    //   iteration m + 1 : warp 0 loads the status words of the super-chunks before it at the
    //                     start, and at the END (after its own chunks) resolves the carry-in and
    //                     derives the governing line of every chunk -> ccarry[m & 1];
    //   iteration m + 2 : every warp folds its share of pend[m % 3] after its second chunk.
    auto resolve_carry = [&](uint32_t sc_prev, uint32_t b, uint32_t cb, unsigned long long pf) {
        uint32_t cnt = C.pend_cnt[b];
        if (cnt > (uint32_t)PENDCAP) {
            if (lane == 0) P.tab.counters[KX_C_PEND_OVERFLOW] = 1u;  // host falls back to v2
            cnt = PENDCAP;
        }
        if (cnt > 0 || C.agg_none[b]) {
            // look-back; first window from the prefetched status words
            unsigned long long carry = 0;
            {
                long long top = (long long)sc_prev - 1;
                unsigned long long sv = pf;
                bool fresh = false;
                for (;;) {
                    const long long idx = top - (long long)lane;
                    if (fresh || idx < 0) sv = idx >= 0 ? ld_volatile_u64(&P.sc_state[idx]) : (ST_PREFIX | P.carry_in);
                    const unsigned long long stt = sv & ST_MASK;
                    const uint32_t pm = __ballot_sync(0xffffffffu, stt == ST_PREFIX);
                    const uint32_t zm = __ballot_sync(0xffffffffu, stt == 0ull);
                    if (pm) {
                        const uint32_t f = (uint32_t)__ffs((int)pm) - 1u;
                        if ((zm & ((1u << f) - 1u)) == 0u) { carry = __shfl_sync(0xffffffffu, sv, (int)f) & ~ST_MASK; break; }
                    } else if (zm == 0u) {
                        top -= 32;
                        fresh = true;
                        continue;
                    }
                    if (fresh) __nanosleep(300);
                    fresh = true;
                }
            }
            if (C.agg_none[b] && lane == 0) st_volatile_u64(&P.sc_state[sc_prev], ST_PREFIX | carry);
            // governing line of the head lines of chunk `lane`: the last top-level line of the
            // chunks before it in the super-chunk, else the super-chunk's carry-in
            const uint32_t x = C.cstate[b][lane];
            const uint32_t tmk = __ballot_sync(0xffffffffu, (x & LS_TOP) != 0u);
            const uint32_t lower = tmk & lt_mask;
            const uint32_t f = lower ? 31u - (uint32_t)__clz((int)lower) : 0u;
            const uint32_t xf = __shfl_sync(0xffffffffu, x, (int)f);
            ChunkCarry cc;
            if (lower) {
                cc.anchor = P.base + (unsigned long long)sc_prev * SCB + f * CW + (xf & 0xfffu);
                cc.key_hi = ((xf >> 12) & 0xffffu) << 16;
                cc.valid = (xf & LS_VOK) ? 1u : 0u;
            } else {
                cc.anchor = carry & CV_ANCHOR_MASK;
                cc.key_hi = ((uint32_t)(carry >> 44) & 0xffffu) << 16;
                cc.valid = ((carry & CV_HAS_TOP) && (carry & CV_VOK)) ? 1u : 0u;
            }
            // same pruning as in the per-line pass: a governing line that is not the first of its
            // vendor id cannot produce a hit
            if (cc.valid && P.tab.vendor_first[cc.key_hi >> 16] < cc.anchor) cc.valid = 0u;
            C.ccarry[cb][lane] = cc;
        }
        if (lane == 0) {
            C.fold_cnt[cb] = cnt;
            C.pend_cnt[b] = 0;  // buffer b is parked into again two iterations from now
        }
        __syncwarp();
    };
    auto fold_share = [&](uint32_t sc_prev, uint32_t b, uint32_t cb) {
        const uint32_t cnt = C.fold_cnt[cb];
        const unsigned long long sbase = P.base + (unsigned long long)sc_prev * SCB + 1ull;  // entries hold the newline position
        for (uint32_t i = w * 32u + lane; i < cnt; i += NT) {
            const uint32_t e = C.pend[b][i];
            const ChunkCarry cc = C.ccarry[cb][(e & 0xffffu) >> 11];
            if (cc.valid) table_fold(P.tab, cc.key_hi | (e >> 16), sbase + (e & 0xffffu), cc.anchor);
        }
    };

    uint32_t phase_bits = 0, q = 0;
    uint32_t sc_m1 = 0, sc_m2 = 0;  // super-chunks of the previous two iterations of this CTA
    unsigned long long pf = 0;      // warp 0: status words of the super-chunks before sc_m1
    uint32_t k = 0;
    for (;; k++) {
        // The iteration boundary is a SPLIT barrier (mbarrier): a warp arrives when it has parsed
        // its chunks of iteration k and only waits for the others' arrival where iteration k + 1
        // first touches CTA-shared buffers (parking / folding), i.e. after the newline masks,
        // classification, scan and publication of its first chunk.
        const uint32_t sc = C.sc_q[k % 3u];
        if (sc >= P.num_sc) break;
        uint32_t sc_next = 0;
        const uint32_t pb = k % 3u;
        auto iter_sync = [&]() {
            if (k >= 1u) mbar_wait_suspend(&C.iter_bar, (k - 1u) & 1u);  // everybody finished iteration k - 1
            if (threadIdx.x == 0) C.sc_q[(k + 2u) % 3u] = atomicAdd(&P.tab.counters[KX_C_TICKET], 1u);  // ticket of iteration k + 2
            sc_next = C.sc_q[(k + 1u) % 3u];
        };
        if (w == 0 && k >= 1u) {
            const long long idx = (long long)sc_m1 - 1 - (long long)lane;
            pf = idx >= 0 ? ld_volatile_u64(&P.sc_state[idx]) : 0ull;
        }
        const uint32_t nsc = P.num_chunks - sc * SCC < (uint32_t)SCC ? P.num_chunks - sc * SCC : (uint32_t)SCC;
        const unsigned long long sc_base = P.base + (unsigned long long)sc * SCB;

        for (uint32_t j = 0; j < (uint32_t)CPW; j++, q++) {
            const uint32_t c = wl + (uint32_t)WARPS * j;  // chunk index inside the super-chunk
            const int s = (int)(q & 1u);
            if (c < nsc) {
                const uint32_t g = sc * SCC + c;
                const uint8_t *st = S.stage[s];
                const uint32_t cpos = c * CW;  // position of st[0] inside the super-chunk
                const unsigned long long cbase = sc_base + cpos;
                uint32_t n_rel = CW + 1;  // line starts at p < n_rel are real (p == CW: first byte of the next chunk)
                if (chunk_tma_ok(g)) {
                    mbar_wait_suspend(&S.bar[s], (phase_bits >> s) & 1u);
                    phase_bits ^= 1u << s;
                } else {
                    // ragged tail of the text: bounded loads, zero fill
                    const unsigned long long chunk_start = (unsigned long long)g * CW;
                    const unsigned long long remain = P.n - chunk_start;
                    n_rel = remain < (unsigned long long)CW ? (uint32_t)remain : (uint32_t)CW + (remain > (unsigned long long)CW);
                    for (int cc = (int)lane; cc < STG_BYTES / 16; cc += 32) {
                        const unsigned long long q0 = chunk_start + 16ull * (unsigned)cc;
                        uint4 v;
                        if (q0 + 16 <= P.n) {
                            v = *reinterpret_cast<const uint4 *>(P.text + q0);
                        } else {
                            uint8_t tmp[16];
#pragma unroll
                            for (int b = 0; b < 16; b++) tmp[b] = q0 + b < P.n ? P.text[q0 + b] : (uint8_t)0;
                            v = *reinterpret_cast<uint4 *>(tmp);
                        }
                        *reinterpret_cast<uint4 *>(S.stage[s] + 16 * cc) = v;
                    }
                    __syncwarp();
                }

                // ------------------------------------------------------ newline masks
                // lane owns bytes [32*lane, 32*lane+32) of each KiB half; the two 16-byte pieces are
                // read in a lane-dependent order so that every LDS.128 phase hits all 32 banks.
                const uint32_t swz = (lane >> 2) & 1u;
                uint32_t kh[2], th[2];  // kept line starts / top-level line starts, bit b: newline at byte b of the lane's window
                uint32_t rawnl = 0;
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const uint32_t o = (uint32_t)h * HALF + lane * 32u;
                    const uint4 va = *reinterpret_cast<const uint4 *>(st + o + 16u * swz);
                    const uint4 vb = *reinterpret_cast<const uint4 *>(st + o + 16u * (swz ^ 1u));
                    const uint32_t ma = nl_mask16(va, k7f, k0a, k80), mb = nl_mask16(vb, k7f, k0a, k80);
                    uint32_t mm = swz ? (mb | (ma << 16)) : (ma | (mb << 16));
                    rawnl |= mm;
                    // a line start at o + 1 + b is real only below n_rel (ragged last chunk)
                    if (n_rel <= (uint32_t)CW) mm &= n_rel > o + 1u ? (n_rel - o - 1u >= 32u ? 0xffffffffu : ((1u << (n_rel - o - 1u)) - 1u)) : 0u;
                    // classify the line that starts after each newline by its first two bytes:
                    //   neither '#' nor '\t': top-level line -- ends the vendor block
                    //     (device_plugin.go:229-236), the only kind locateVendor can match (:265)
                    //   "\t" + non-tab: device line candidate (:237); "\t\t" subsystem, '#' comment: dropped
                    uint32_t km = 0, tm = 0;
                    const uint8_t *lp = st + o + 1u;
                    while (mm) {
                        const uint32_t b = (uint32_t)__ffs((int)mm) - 1u;
                        const uint32_t bit = mm & (0u - mm);
                        mm ^= bit;
                        const uint32_t c0 = lp[b], c1 = lp[b + 1u];
                        // top  = c0 != '\t' && c0 != '#'        -> tm |= bit
                        // cand = c0 == '\t' && c1 != '\t'       -> km |= bit (with top)
                        // written with predicates: 6 instructions instead of the 11 the compiler emits
                        asm("{\n\t.reg .pred p0, pt, pc, pk;\n\t"
                            "setp.eq.u32 p0, %2, 9;\n\t"
                            "setp.ne.and.u32 pt, %2, 35, !p0;\n\t"
                            "setp.ne.and.u32 pc, %3, 9, p0;\n\t"
                            "or.pred pk, pt, pc;\n\t"
                            "@pt or.b32 %0, %0, %4;\n\t"
                            "@pk or.b32 %1, %1, %4;\n\t}"
                            : "+r"(tm), "+r"(km)
                            : "r"(c0), "r"(c1), "r"(bit));
                    }
                    kh[h] = km;
                    th[h] = tm;
                }
                // the shard starts with a line start at p = 0 (no newline before it)
                uint32_t inj = 0, inj_top = 0;
                if (g == 0) {
                    const uint32_t c0 = st[0], c1 = st[1];
                    const bool real = 0u < n_rel;
                    inj_top = (real && c0 != (uint32_t)'#' && c0 != (uint32_t)'\t') ? 1u : 0u;
                    inj = (inj_top || (real && c0 == (uint32_t)'\t' && c1 != (uint32_t)'\t')) ? 1u : 0u;
                }
                // 2 KiB without a newline may belong to a >= 64 KiB line (bufio.ErrTooLong): raise the
                // hint, the exact cut-off is then computed by trunc_kernel (never for real pci.ids)
                if (n_rel > (uint32_t)CW && __reduce_or_sync(0xffffffffu, rawnl) == 0u && lane == 0)
                    atomicOr(&P.tab.counters[KX_C_LONGLINE_HINT], 1u);
                const uint32_t cnt = (uint32_t)__popc(kh[0]) | ((uint32_t)__popc(kh[1]) << 16);
                uint32_t incl = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= (uint32_t)d) incl += y;
                }
                const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
                const uint32_t tot0 = tot & 0xffffu;
                const uint32_t K = tot0 + (tot >> 16) + inj;  // kept lines of this chunk
                const uint32_t excl = incl - cnt;
                const uint32_t off0 = inj + (excl & 0xffffu), off1 = inj + tot0 + (excl >> 16);

                // last top-level line of the chunk (position + 1, 0 = none)
                uint32_t lt1 = inj_top ? 1u : 0u;
                if (th[0]) lt1 = lane * 32u + 1u + (31u - (uint32_t)__clz((int)th[0])) + 1u;
                if (th[1]) lt1 = (uint32_t)HALF + lane * 32u + 1u + (31u - (uint32_t)__clz((int)th[1])) + 1u;
                lt1 = __reduce_max_sync(0xffffffffu, lt1);

                // publish it to the CTA; the warp that publishes last publishes the super-chunk's
                // aggregate to the grid.
                {
                    uint32_t ls = LS_PUB;
                    if (lt1) {
                        uint32_t v;
                        const bool ok = hex4_swar(lds_u32_unaligned(st, lt1 - 1u), v);
                        ls |= LS_TOP | (ok ? LS_VOK : 0u) | (v << 12) | (lt1 - 1u);
                    }
                    uint32_t old = 0;
                    if (lane == 0) {
                        *reinterpret_cast<volatile uint32_t *>(&C.cstate[pb][c]) = ls;
                        __threadfence_block();
                        old = atomicAdd(&C.pub_cnt[k & 1u], 1u);
                    }
                    old = __shfl_sync(0xffffffffu, old, 0);
                    if (old + 1u == nsc) {
                        __threadfence_block();
                        const uint32_t x = lane < nsc ? *reinterpret_cast<volatile uint32_t *>(&C.cstate[pb][lane]) : 0u;
                        const uint32_t hm = __ballot_sync(0xffffffffu, (x & LS_TOP) != 0u);
                        if (hm) {
                            const uint32_t fl = 31u - (uint32_t)__clz((int)hm);
                            const uint32_t xl = __shfl_sync(0xffffffffu, x, (int)fl);
                            const unsigned long long own = CV_HAS_TOP | ((xl & LS_VOK) ? CV_VOK : 0ull) |
                                                           ((unsigned long long)((xl >> 12) & 0xffffu) << 44) |
                                                           ((sc_base + fl * CW + (xl & 0xfffu)) & CV_ANCHOR_MASK);
                            if (lane == 0) st_volatile_u64(&P.sc_state[sc], ST_PREFIX | own);
                        } else if (lane == 0) {
                            st_volatile_u64(&P.sc_state[sc], ST_NONE);
                        }
                        if (lane == 0) {
                            C.agg_none[pb] = hm ? 0u : 1u;  // read one iteration later
                            C.pub_cnt[k & 1u] = 0;          // next used two iterations from now
                        }
                    }
                }
                if (j == 0u) iter_sync();

                // ------------------------------------------------------ one lane per kept line
                uint32_t cP = P_NONE;  // governing top-level line so far; P_NONE = before the chunk's first one
                for (uint32_t lo = 0; lo < K; lo += LCAP) {
                    // ordered list of the kept line starts [lo, lo + LCAP): entry = position | top << 15
                    if (inj && lane == 0 && lo == 0) S.list[0] = (uint16_t)(inj_top ? 0x8000u : 0u);
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        uint32_t idx = (h == 0 ? off0 : off1) - lo;
                        uint32_t mm = kh[h];
                        const uint32_t tmh = th[h];
                        const uint32_t p0 = (uint32_t)h * HALF + lane * 32u + 1u;  // a line starts one past its newline
                        while (mm) {
                            const uint32_t b = (uint32_t)__ffs((int)mm) - 1u;
                            const uint32_t bit = mm & (0u - mm);
                            mm ^= bit;
                            if (idx < (uint32_t)LCAP) S.list[idx] = (uint16_t)((p0 + b) | ((tmh & bit) ? 0x8000u : 0u));
                            idx++;
                        }
                    }
                    __syncwarp();
                    const uint32_t kept = K - lo < (uint32_t)LCAP ? K - lo : (uint32_t)LCAP;
                    for (uint32_t r = 0; r < kept; r += 32u) {
                        const uint32_t i = r + lane;
                        const bool act = i < kept;
                        const uint32_t e = act ? (uint32_t)S.list[i] : 0u;
                        const uint32_t p = e & 0x0fffu;
                        const bool istop = act && (e >> 15);
                        uint32_t val;
                        const bool ok = hex4_swar(lds_u32_unaligned(st, p + (istop ? 0u : 1u)), val);
                        const bool isdev = act && !istop && ok;
                        const unsigned long long line_g = cbase + p;
                        // A top-level line is a candidate vendor anchor; only the FIRST line with this
                        // prefix counts (:265).  If an earlier one is already known, this block can
                        // never produce a hit (a hit needs min_anchor == vendor_first): its device
                        // lines are dropped right here instead of being folded into the table.
                        bool alive = ok;
                        if (istop && ok) {
                            const unsigned long long vf = P.tab.vendor_first[val];
                            if (line_g < vf) atomicMin(&P.tab.vendor_first[val], line_g);
                            alive = line_g <= vf;
                        }
                        const uint32_t myP = (alive ? 0x80000000u : 0u) | (val << 15) | p;
                        const uint32_t tm = __ballot_sync(0xffffffffu, istop);
                        const uint32_t prev = tm & lt_mask;
                        const uint32_t g_src = __shfl_sync(0xffffffffu, myP, prev ? 31 - __clz((int)prev) : 0);
                        const uint32_t gov = prev ? g_src : cP;
                        const bool was_head = cP == P_NONE;
                        if (tm) cP = __shfl_sync(0xffffffffu, myP, 31 - __clz((int)tm));
                        if (was_head) {
                            // device lines before the chunk's first top-level line are governed by an
                            // earlier chunk: park them (resolve_prev folds them one iteration later)
                            const bool headdev = isdev && gov == P_NONE;
                            if (headdev && p == 0u) {
                                // the shard's first line: governed by the shard's carry-in, which is known
                                if ((P.carry_in & CV_HAS_TOP) && (P.carry_in & CV_VOK))
                                    table_fold(P.tab, (((uint32_t)(P.carry_in >> 44) & 0xffffu) << 16) | val, line_g,
                                               P.carry_in & CV_ANCHOR_MASK);
                            }
                            const bool park = headdev && p != 0u;
                            const uint32_t hm = __ballot_sync(0xffffffffu, park);
                            if (hm) {
                                uint32_t basei = 0;
                                if (lane == 0) basei = atomicAdd(&C.pend_cnt[pb], (uint32_t)__popc(hm));
                                basei = __shfl_sync(0xffffffffu, basei, 0);
                                const uint32_t idx = basei + (uint32_t)__popc(hm & lt_mask);
                                if (park && idx < (uint32_t)PENDCAP) C.pend[pb][idx] = (val << 16) | (cpos + p - 1u);
                            }
                        }
                        if (isdev && gov != P_NONE && (gov >> 31))
                            table_fold(P.tab, (((gov >> 15) & 0xffffu) << 16) | val, line_g, cbase + (gov & 0x7fffu));
                    }
                    __syncwarp();
                }
            } else if (j == 0u) {
                iter_sync();
            }

            // prefetch the chunk two steps ahead of this warp into the stage just released
            if (lane == 0) {
                if (j + 2u < (uint32_t)CPW) issue(sc, c + 2u * WARPS, s);
                else issue(sc_next, c + 2u * WARPS - (uint32_t)SCC, s);
            }
            if (j == 1u && k >= 2u) fold_share(sc_m2, (k - 2u) % 3u, k & 1u);
        }
        if (w == 0 && k >= 1u) resolve_carry(sc_m1, (k - 1u) % 3u, (k - 1u) & 1u, pf);
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&C.iter_bar)) : "memory");
        sc_m2 = sc_m1;
        sc_m1 = sc;
    }
    // drain: k iterations were run
    if (k >= 1u) mbar_wait_suspend(&C.iter_bar, (k - 1u) & 1u);
    if (k >= 2u) fold_share(sc_m2, (k - 2u) % 3u, k & 1u);
    if (k >= 1u) {
        if (w == 0) {
            const long long idx = (long long)sc_m1 - 1 - (long long)lane;
            resolve_carry(sc_m1, (k - 1u) % 3u, (k - 1u) & 1u, idx >= 0 ? ld_volatile_u64(&P.sc_state[idx]) : 0ull);
        }
        __syncthreads();
        fold_share(sc_m1, (k - 1u) % 3u, (k - 1u) & 1u);
    }
}

}  // namespace kxparse3
