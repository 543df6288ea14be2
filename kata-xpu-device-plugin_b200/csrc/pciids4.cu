// pciids4.cu -- parse kernel v4: alive-first, no waiting, nothing parked, nothing listed.
//
// What the ncu profile of v3 showed (profiles/r01_parse_v3h_ncu_summary.txt): the kernel is
// bound by instruction issue (0.33 warp-instructions per byte), and 45 % of them order the kept
// lines into a list, walk it with one lane per line and park the lines whose vendor line sits
// in an earlier chunk -- although, with first-occurrence-wins (device_plugin.go:265), a device
// line only matters when the vendor line that governs it is the FIRST line with that id.
// v4 decides that first and touches a device line only when its block is alive:
//   * the text is cut into RANGES of 128 KiB handed out by ticket; a CTA (8 warps) walks its
//     range as super-chunks of 8 consecutive 2 KiB chunks, one chunk per warp, staged by 1-D
//     TMA bulk copies into a private 3-stage ring;
//   * PHASE A (iteration k): newline masks and line classes as in v3, then the TOP-LEVEL lines
//     only: hex prefix, vendor_first check/update, alive bit.  Every lane keeps the last
//     top-level line of its 32-byte windows; two ballots + shuffles give each window its
//     governing line inside the chunk.  Device lines under an alive line are folded on the spot
//     (rare: once per vendor id); under a dead one they are dropped without being parsed.  The
//     chunk's last top-level line goes to shared memory;
//   * PHASE B (iteration k + 1, after phase A of the next super-chunk, behind one mbarrier
//     arrive/wait pair with a whole phase of slack): the lines in front of the chunk's first
//     top-level line ("head" lines, two bit masks in registers) take their governing line from
//     the chunks before them in the super-chunk, else from the carry every warp keeps along the
//     range, and are folded only if that line is alive;
//   * at the start of a range the carry is not known (the range before belongs to another CTA):
//     such chunks are DEFERRED -- 4 bytes in a list -- and a small second kernel resolves them
//     from the per-range status words once everything is published (and re-parses the chunk in
//     the rare case that its governing line is alive).  The main kernel never waits for
//     another CTA.
#pragma once
#include "common.cuh"
#include "pciids2.cu"  // nl_mask16, hex4_swar, table_fold, TMA/mbarrier helpers, CV_*/ST_* encodings
#include "pciids3.cu"  // LS_* chunk status encoding
#include "table.cuh"

namespace kxparse4 {

using namespace kxparse2;

constexpr int STAGES4 = 3;
constexpr int SCC4 = WARPS;        // chunks per super-chunk: one per warp
constexpr int SCB4 = SCC4 * CW;    // 16384 bytes
constexpr int RSC = 8;             // super-chunks per range
constexpr int RCH = RSC * SCC4;    // chunks per range (64 = 128 KiB)
constexpr int RES_WARPS = 8;

struct WarpSmem4 {
    alignas(16) uint8_t stage[STAGES4][STG_BYTES];
    alignas(8) unsigned long long bar[STAGES4];
};

struct CtaSmem4 {
    WarpSmem4 w[WARPS];
    uint32_t cstate[4][SCC4];  // see LS_* (pciids3.cu); [k % 4] = iteration k
    uint32_t rq[4];            // range tickets, [j % 4] = j-th range of this CTA
    alignas(8) unsigned long long it_bar[2];  // mbarrier [k & 1]: phase A of iteration k done by all warps
};

struct Params4 {
    const uint8_t *text;
    unsigned long long n, base;
    uint32_t num_chunks, num_sc, num_ranges;
    uint32_t tma_limit;               // chunks [0, tma_limit) can be staged with one bulk copy of STG_BYTES
    unsigned long long *range_state;  // [num_ranges]
    uint32_t *deferred;               // [num_chunks] chunk indices, count in counters[KX_C_DEFER]
    KxTableDev tab;
    unsigned long long carry_in;
};

// ---- shared memory through explicit 32-bit shared-window addresses --------------------------
// The compiler re-derives the shared window base (S2R SR_CgaCtaId + LEA) at every use of a
// generic pointer into dynamic shared memory; the hot loop therefore keeps ONE base address in a
// register and goes through ld/st.shared with integer offsets.
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds8(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
// four bytes at an arbitrary shared-memory byte address
__device__ __forceinline__ uint32_t lds32_unaligned(uint32_t a) {
    const uint32_t al = a & ~3u;
    return __funnelshift_r(lds32(al), lds32(al + 4u), (a & 3u) * 8u);
}
__device__ __forceinline__ void mbar_expect_tx_a(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ bool mbar_try_a(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
    return ok != 0u;
}
__device__ __forceinline__ void tma_load_a(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, unsigned long long pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar), "l"(pol)
                 : "memory");
}

// stage chunk g with bounded loads and zero fill (ragged tail of the text / the resolve kernel);
// returns n_rel: line starts at p < n_rel are real
__device__ __forceinline__ uint32_t stage_chunk_manual(const uint8_t *text, unsigned long long n, uint32_t g, uint32_t lane,
                                                       uint8_t *dst) {
    const unsigned long long chunk_start = (unsigned long long)g * CW;
    const unsigned long long remain = n - chunk_start;
    const uint32_t n_rel = remain < (unsigned long long)CW ? (uint32_t)remain : (uint32_t)CW + (remain > (unsigned long long)CW);
    for (int cc = (int)lane; cc < STG_BYTES / 16; cc += 32) {
        const unsigned long long q0 = chunk_start + 16ull * (unsigned)cc;
        uint4 v;
        if (q0 + 16 <= n) {
            v = *reinterpret_cast<const uint4 *>(text + q0);
        } else {
            uint8_t tmp[16];
#pragma unroll
            for (int b = 0; b < 16; b++) tmp[b] = q0 + b < n ? text[q0 + b] : (uint8_t)0;
            v = *reinterpret_cast<uint4 *>(tmp);
        }
        *reinterpret_cast<uint4 *>(dst + 16 * cc) = v;
    }
    __syncwarp();
    return n_rel;
}

// Newline masks and line classes of the chunk staged at shared address st.  Lane owns bytes
// [32*lane, 32*lane+32) of each KiB half; the two 16-byte pieces are read in a lane-dependent
// order so that every LDS.128 phase hits all banks.  kh / th: kept / top-level line starts,
// bit b = the line that starts after a newline at byte b of the lane's window.
__device__ __forceinline__ void chunk_masks(uint32_t st, uint32_t lane, uint32_t n_rel, uint32_t k7f, uint32_t k0a, uint32_t k80,
                                            uint32_t (&kh)[2], uint32_t (&th)[2], uint32_t &rawnl) {
    const uint32_t swz = (lane >> 2) & 1u;
    rawnl = 0;
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const uint32_t o = (uint32_t)h * HALF + lane * 32u;
        const uint4 va = lds128(st + o + 16u * swz);
        const uint4 vb = lds128(st + o + 16u * (swz ^ 1u));
        const uint32_t ma = nl_mask16(va, k7f, k0a, k80), mb = nl_mask16(vb, k7f, k0a, k80);
        uint32_t mm = swz ? (mb | (ma << 16)) : (ma | (mb << 16));
        rawnl |= mm;
        // a line start at o + 1 + b is real only below n_rel (ragged last chunk)
        if (n_rel <= (uint32_t)CW) mm &= n_rel > o + 1u ? (n_rel - o - 1u >= 32u ? 0xffffffffu : ((1u << (n_rel - o - 1u)) - 1u)) : 0u;
        // class of the line that starts after each newline, by its first two bytes:
        //   neither '#' nor '\t': top-level line -- ends the vendor block
        //     (device_plugin.go:229-236), the only kind locateVendor can match (:265)
        //   "\t" + non-tab: device line candidate (:237); "\t\t" subsystem, '#' comment: dropped
        uint32_t km = 0, tm = 0;
        const uint32_t lp = st + o + 1u;
        while (mm) {
            const uint32_t b = (uint32_t)__ffs((int)mm) - 1u;
            const uint32_t bit = mm & (0u - mm);
            mm ^= bit;
            const uint32_t c0 = lds8(lp + b), c1 = lds8(lp + b + 1u);
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
}

// device lines `m` (bit b: line starts at pbase + b) of the chunk staged at st, all governed by
// the alive top-level line (key_hi, anchor): parse the id, fold.  Per-lane loop: only blocks of
// a first-seen vendor id get here.
__device__ __forceinline__ void fold_lines(const KxTableDev &tab, uint32_t st, unsigned long long cbase, uint32_t m, uint32_t pbase,
                                           uint32_t key_hi, unsigned long long anchor) {
    while (m) {
        const uint32_t b = (uint32_t)__ffs((int)m) - 1u;
        m &= m - 1u;
        const uint32_t p = pbase + b;
        uint32_t dv;
        if (hex4_swar(lds32_unaligned(st + p + 1u), dv)) table_fold(tab, key_hi | dv, cbase + p, anchor);
    }
}

__global__ void __launch_bounds__(NT, 4) parse_kernel_v4(const Params4 P) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    CtaSmem4 &C = *reinterpret_cast<CtaSmem4 *>(smem_raw);
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    // the issue arbiter favours the highest warp id: it gets the earliest chunk
    const uint32_t c = (uint32_t)WARPS - 1u - w;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t c_mask = (1u << c) - 1u;

    if (threadIdx.x == 0) {
        C.rq[0] = atomicAdd(&P.tab.counters[KX_C_TICKET], 1u);
        mbar_init(&C.it_bar[0], WARPS);
        mbar_init(&C.it_bar[1], WARPS);
    }
    if (lane == 0) {
        for (int s = 0; s < STAGES4; s++) mbar_init(&C.w[w].bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // shared-window addresses (see lds128 above)
    const uint32_t sb = smem_u32(smem_raw);
    const uint32_t a_stage0 = sb + (uint32_t)offsetof(CtaSmem4, w) + w * (uint32_t)sizeof(WarpSmem4);  // stage s: + s * STG_BYTES
    const uint32_t a_bar0 = a_stage0 + (uint32_t)offsetof(WarpSmem4, bar);                            // bar s:   + 8 * s
    const uint32_t a_cstate = sb + (uint32_t)offsetof(CtaSmem4, cstate);
    const uint32_t a_rq = sb + (uint32_t)offsetof(CtaSmem4, rq);
    const uint32_t a_itbar = sb + (uint32_t)offsetof(CtaSmem4, it_bar);

    uint32_t k7f = 0x7f7f7f7fu, k0a = 0x0a0a0a0au, k80 = 0x80808080u;
    asm volatile("" : "+r"(k7f), "+r"(k0a), "+r"(k80));

    const unsigned long long pol = l2_evict_first_policy();
    auto issue = [&](uint32_t g, uint32_t s) {  // lane 0: start the TMA copy of chunk g into stage s
        if (g < P.tma_limit) {
            mbar_expect_tx_a(a_bar0 + 8u * s, STG_BYTES);
            tma_load_a(a_stage0 + s * (uint32_t)STG_BYTES, P.text + (unsigned long long)g * CW, STG_BYTES, a_bar0 + 8u * s, pol);
        }
    };

    // three cursors walk the CTA's sequence of ranges: A = the iteration's own super-chunk,
    // B = the one before it (phase B lags one iteration), F = two ahead (TMA prefetch).
    // g / f_g: my chunk of the A / F super-chunk (0xffffffff: no more ranges)
    uint32_t a_rt = lds32(a_rq), a_i = 0, a_j = 0;
    uint32_t g = a_rt < P.num_ranges ? a_rt * RCH + c : 0xffffffffu;
    uint32_t f_g = g == 0xffffffffu ? g : g + 2u * SCC4, f_i = 2, f_j = 0;
    if (lane == 0 && g != 0xffffffffu) {
        issue(g, 0);
        issue(g + SCC4, 1);
    }

    // carry along the range (every warp keeps its own, identical copy): the governing line at
    // the start of the next super-chunk as a chunk status word (0 = not known) plus the chunk
    // that holds the line (0xffffffff = the shard's carry-in)
    uint32_t rc_x = 0, rc_g = 0, prev_rt = 0;
    // carried from phase A of an iteration into its phase B
    uint32_t hw0 = 0, hw1 = 0, b_g = 0, b_i = 0, b_rt = 0;
    bool need_b = false;

    // PHASE B: head lines hw0 / hw1 of my chunk b_g (staged at st), then the carry
    auto phase_b = [&](uint32_t st, uint32_t cb) {
        if (b_i == 0u) {
            if (b_rt == 0u) {
                rc_x = LS_PUB | (((P.carry_in & CV_HAS_TOP) && (P.carry_in & CV_VOK)) ? (LS_TOP | LS_VOK) : 0u);
                rc_g = 0xffffffffu;
            } else if (b_rt != prev_rt + 1u) {
                rc_x = 0u;  // the range before this one belongs to another CTA
            }
            prev_rt = b_rt;
        }
        const uint32_t x = lane < (uint32_t)SCC4 ? lds32(a_cstate + cb * (4u * SCC4) + 4u * lane) : 0u;
        const uint32_t tmk_all = __ballot_sync(0xffffffffu, (x & LS_TOP) != 0u);
        if (need_b) {
            uint32_t gx = rc_x, gg = rc_g;
            const uint32_t tmk = tmk_all & c_mask;  // chunks before mine
            if (tmk) {
                const uint32_t f = 31u - (uint32_t)__clz((int)tmk);
                gx = __shfl_sync(0xffffffffu, x, (int)f);
                gg = b_g - c + f;
            }
            if (gx == 0u) {
                if (lane == 0) P.deferred[atomicAdd(&P.tab.counters[KX_C_DEFER], 1u)] = b_g;
            } else if (gx & LS_VOK) {
                uint32_t key_hi = ((gx >> 12) & 0xffffu) << 16;
                unsigned long long anchor = P.base + (unsigned long long)gg * CW + (gx & 0xfffu);
                if (gg == 0xffffffffu) {
                    key_hi = ((uint32_t)(P.carry_in >> 44) & 0xffffu) << 16;
                    anchor = P.carry_in & CV_ANCHOR_MASK;
                }
                // still the first line of its id?
                if ((hw0 | hw1) != 0u && P.tab.vendor_first[key_hi >> 16] >= anchor) {
                    const unsigned long long cbase = P.base + (unsigned long long)b_g * CW;
                    fold_lines(P.tab, st, cbase, hw0, lane * 32u + 1u, key_hi, anchor);
                    fold_lines(P.tab, st, cbase, hw1, (uint32_t)HALF + lane * 32u + 1u, key_hi, anchor);
                }
            }
        }
        if (tmk_all) {
            const uint32_t fl = 31u - (uint32_t)__clz((int)tmk_all);
            rc_x = __shfl_sync(0xffffffffu, x, (int)fl);
            rc_g = b_g - c + fl;
        }
        // end of the range: its inclusive carry for the resolve kernel
        if ((b_i == (uint32_t)RSC - 1u || b_g - c + (uint32_t)SCC4 >= P.num_chunks) && threadIdx.x == (uint32_t)(NT - 32)) {
            unsigned long long v = ST_NONE;
            if (rc_x != 0u) {
                if (rc_g == 0xffffffffu)
                    v = ST_PREFIX | P.carry_in;
                else
                    v = ST_PREFIX | CV_HAS_TOP | ((rc_x & LS_VOK) ? CV_VOK : 0ull) | ((unsigned long long)((rc_x >> 12) & 0xffffu) << 44) |
                        ((P.base + (unsigned long long)rc_g * CW + (rc_x & 0xfffu)) & CV_ANCHOR_MASK);
            }
            P.range_state[b_rt] = v;
        }
    };
    // wait for phase A of iteration kk of all warps.  try_wait + nanosleep: the hardware
    // suspend variant wakes on every barrier event of the CTA (14 spins per wait measured)
    auto wait_iter = [&](uint32_t kk) {
        const uint32_t bar = a_itbar + 8u * (kk & 1u), par = (kk >> 1) & 1u;
        while (!mbar_try_a(bar, par)) __nanosleep(256);
    };

    uint32_t phase_bits = 0;
    uint32_t s = 0, sp = 2;  // stage of this iteration / of the previous one (= of the one two ahead)
    uint32_t k = 0;
    for (;; k++) {
        if (g == 0xffffffffu) break;
        uint32_t tk = 0;
        const bool draw = threadIdx.x == 0 && a_i == 0u;
        if (draw) tk = atomicAdd(&P.tab.counters[KX_C_TICKET], 1u);  // next range of this CTA
        uint32_t n_hw0 = 0, n_hw1 = 0;
        bool n_need = false;
        const uint32_t st = a_stage0 + s * (uint32_t)STG_BYTES;

        // ---------------------------------------------------------------- PHASE A(k)
        if (g < P.num_chunks) {
            const unsigned long long cbase = P.base + (unsigned long long)g * CW;
            uint32_t n_rel = CW + 1;  // line starts at p < n_rel are real (p == CW: first byte of the next chunk)
            if (g < P.tma_limit) {
                const uint32_t bar = a_bar0 + 8u * s, par = (phase_bits >> s) & 1u;
                while (!mbar_try_a(bar, par)) {
                }
                phase_bits ^= 1u << s;
            } else {
                n_rel = stage_chunk_manual(P.text, P.n, g, lane, C.w[w].stage[s]);
            }

            uint32_t kh[2], th[2], rawnl;
            chunk_masks(st, lane, n_rel, k7f, k0a, k80, kh, th, rawnl);

            // the shard starts with a line start at p = 0 (no newline before it)
            uint32_t base_info = P_NONE;  // top-level line in front of the lane windows (only that one)
            if (g == 0u && n_rel > 0u) {
                const uint32_t c0 = lds8(st), c1 = lds8(st + 1u);
                if (c0 != (uint32_t)'#' && c0 != (uint32_t)'\t') {
                    uint32_t val;
                    const bool ok = hex4_swar(lds32_unaligned(st), val);
                    bool alive = ok;
                    if (ok) {
                        const unsigned long long vf = P.tab.vendor_first[val];
                        if (lane == 0 && cbase < vf) atomicMin(&P.tab.vendor_first[val], cbase);
                        alive = cbase <= vf;
                    }
                    base_info = (alive ? 0x80000000u : 0u) | ((ok ? val : 0u) << 15);
                } else if (c0 == (uint32_t)'\t' && c1 != (uint32_t)'\t') {
                    // device line at the very start: governed by the shard's carry-in, which is known
                    uint32_t dv;
                    if (lane == 0 && (P.carry_in & CV_HAS_TOP) && (P.carry_in & CV_VOK) && hex4_swar(lds32_unaligned(st + 1u), dv))
                        table_fold(P.tab, (((uint32_t)(P.carry_in >> 44) & 0xffffu) << 16) | dv, cbase, P.carry_in & CV_ANCHOR_MASK);
                }
            }

            // top-level lines, by the lane that owns them: a candidate vendor anchor; only the FIRST
            // line with this prefix counts (:265).  If an earlier one is already known, this block
            // can never produce a hit (a hit needs min_anchor == vendor_first): it is dead.
            uint32_t linfo0 = P_NONE, linfo1 = P_NONE;  // last top-level line of my windows: alive<<31 | vendor<<15 | position
            bool any_alive = base_info != P_NONE;
            {
                uint32_t t0 = th[0], t1 = th[1];
                while (t0 | t1) {
                    const bool second = t0 == 0u;
                    const uint32_t tmv = second ? t1 : t0;
                    const uint32_t bit = tmv & (0u - tmv);
                    const uint32_t rest = tmv ^ bit;
                    if (second) t1 = rest; else t0 = rest;
                    const uint32_t pbase = (second ? (uint32_t)HALF : 0u) + lane * 32u + 1u;
                    const uint32_t p = pbase + (31u - (uint32_t)__clz((int)bit));
                    uint32_t val;
                    const bool ok = hex4_swar(lds32_unaligned(st + p), val);
                    const unsigned long long line_g = cbase + p;
                    bool alive = ok;
                    if (ok) {
                        const unsigned long long vf = P.tab.vendor_first[val];
                        if (line_g < vf) atomicMin(&P.tab.vendor_first[val], line_g);
                        alive = line_g <= vf;
                    }
                    if (alive) {
                        // device lines of my window between this line and the next top-level line
                        const uint32_t nxt = rest & (0u - rest);
                        const uint32_t seg = (second ? kh[1] & ~th[1] : kh[0] & ~th[0]) & ~(bit | (bit - 1u)) & (nxt ? nxt - 1u : 0xffffffffu);
                        fold_lines(P.tab, st, cbase, seg, pbase, val << 16, line_g);
                        any_alive = true;
                    }
                    const uint32_t info = (alive ? 0x80000000u : 0u) | ((ok ? val : 0u) << 15) | p;
                    if (second) linfo1 = info; else linfo0 = info;
                }
            }
            const uint32_t bal0 = __ballot_sync(0xffffffffu, th[0] != 0u);
            const uint32_t bal1 = __ballot_sync(0xffffffffu, th[1] != 0u);
            // device lines in front of a window's first top-level line
            const uint32_t pre0 = kh[0] & ~th[0] & (th[0] ? (th[0] & (0u - th[0])) - 1u : 0xffffffffu);
            const uint32_t pre1 = kh[1] & ~th[1] & (th[1] ? (th[1] & (0u - th[1])) - 1u : 0xffffffffu);
            uint32_t last1;  // the chunk's last top-level line
            if (!__any_sync(0xffffffffu, any_alive)) {
                // common case: nothing alive in this chunk.  Windows behind the chunk's first
                // top-level line are dead, the ones in front of it are head lines (phase B).
                const uint32_t bl = bal1 ? bal1 : bal0;
                last1 = __shfl_sync(0xffffffffu, bal1 ? linfo1 : linfo0, bl ? 31 - __clz((int)bl) : 0);
                if (bl == 0u) last1 = P_NONE;
                n_hw0 = (bal0 & lt_mask) == 0u ? pre0 : 0u;
                n_hw1 = (bal0 == 0u && (bal1 & lt_mask) == 0u) ? pre1 : 0u;
            } else {
                // governing line of every window inside the chunk (P_NONE: none, head lines)
                const uint32_t s0 = bal0 & lt_mask, s1 = bal1 & lt_mask;
                const uint32_t x0 = __shfl_sync(0xffffffffu, linfo0, s0 ? 31 - __clz((int)s0) : 0);
                const uint32_t l0 = __shfl_sync(0xffffffffu, linfo0, bal0 ? 31 - __clz((int)bal0) : 0);
                const uint32_t x1 = __shfl_sync(0xffffffffu, linfo1, s1 ? 31 - __clz((int)s1) : 0);
                const uint32_t l1 = __shfl_sync(0xffffffffu, linfo1, bal1 ? 31 - __clz((int)bal1) : 0);
                const uint32_t last0 = bal0 ? l0 : base_info;
                const uint32_t cin0 = s0 ? x0 : base_info;
                const uint32_t cin1 = s1 ? x1 : last0;
                last1 = bal1 ? l1 : last0;
                // governed by an alive line of an earlier window of this chunk: fold now
                if (cin0 != P_NONE && (cin0 >> 31))
                    fold_lines(P.tab, st, cbase, pre0, lane * 32u + 1u, ((cin0 >> 15) & 0xffffu) << 16, cbase + (cin0 & 0x7fffu));
                if (cin1 != P_NONE && (cin1 >> 31))
                    fold_lines(P.tab, st, cbase, pre1, (uint32_t)HALF + lane * 32u + 1u, ((cin1 >> 15) & 0xffffu) << 16, cbase + (cin1 & 0x7fffu));
                n_hw0 = cin0 == P_NONE ? pre0 : 0u;
                n_hw1 = cin1 == P_NONE ? pre1 : 0u;
            }
            n_need = __any_sync(0xffffffffu, (n_hw0 | n_hw1) != 0u);
            // publish the last top-level line to the CTA (read after the barrier, in phase B)
            if (lane == 0) {
                uint32_t ls = LS_PUB;
                if (last1 != P_NONE) ls |= LS_TOP | ((last1 >> 31) ? LS_VOK : 0u) | (((last1 >> 15) & 0xffffu) << 12) | (last1 & 0xfffu);
                sts32(a_cstate + (k & 3u) * (4u * SCC4) + 4u * c, ls);
            }
            // 2 KiB without a newline may belong to a >= 64 KiB line (bufio.ErrTooLong): raise the
            // hint, the exact cut-off is then computed by trunc_kernel (never for real pci.ids)
            if (!n_need && (bal0 | bal1) == 0u && n_rel > (uint32_t)CW && __reduce_or_sync(0xffffffffu, rawnl) == 0u && lane == 0)
                atomicOr(&P.tab.counters[KX_C_LONGLINE_HINT], 1u);
        } else if (lane == 0) {
            sts32(a_cstate + (k & 3u) * (4u * SCC4) + 4u * c, LS_PUB);  // beyond the text
        }
        if (draw) sts32(a_rq + 4u * ((a_j + 1u) & 3u), tk);
        __syncwarp();
        if (lane == 0) mbar_arrive_a(a_itbar + 8u * (k & 1u));

        // ---------------------------------------------------------------- PHASE B(k - 1)
        const uint32_t stp = a_stage0 + sp * (uint32_t)STG_BYTES;
        if (k >= 1u) {
            wait_iter(k - 1u);
            phase_b(stp, (k - 1u) & 3u);
            __syncwarp();
        }
        hw0 = n_hw0;
        hw1 = n_hw1;
        need_b = n_need;
        b_g = g;
        b_i = a_i;
        b_rt = a_rt;
        // the stage of iteration k - 1 is free: prefetch my chunk of iteration k + 2 into it
        if (lane == 0 && f_g != 0xffffffffu) issue(f_g, sp);
        // advance the cursors
        f_g += SCC4;
        if (++f_i == (uint32_t)RSC) {
            f_i = 0;
            const uint32_t rt = lds32(a_rq + 4u * (++f_j & 3u));
            f_g = rt < P.num_ranges ? rt * RCH + c : 0xffffffffu;
        }
        g += SCC4;
        if (++a_i == (uint32_t)RSC) {
            a_i = 0;
            a_rt = lds32(a_rq + 4u * (++a_j & 3u));
            g = a_rt < P.num_ranges ? a_rt * RCH + c : 0xffffffffu;
        }
        sp = s;
        s = s == 2u ? 0u : s + 1u;
    }
    // drain: k iterations were run
    if (k >= 1u) {
        wait_iter(k - 1u);
        phase_b(a_stage0 + sp * (uint32_t)STG_BYTES, (k - 1u) & 3u);
    }
}

// Second kernel: the chunks whose governing line was not known to the CTA that parsed them
// (start of a range).  One thread per chunk walks back over the range status words (all
// published now); the governing line is dead for all but the first copy of a vendor block, and
// only then the warp stages the chunk again and folds its head lines.
__global__ void __launch_bounds__(RES_WARPS * 32) resolve_deferred_kernel(const Params4 P) {
    __shared__ __align__(16) uint8_t stg[RES_WARPS][STG_BYTES];
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t k7f = 0x7f7f7f7fu, k0a = 0x0a0a0a0au, k80 = 0x80808080u;
    asm volatile("" : "+r"(k7f), "+r"(k0a), "+r"(k80));
    const uint32_t n_def = P.tab.counters[KX_C_DEFER];
    const uint32_t st = smem_u32(stg[w]);
    for (uint32_t i0 = blockIdx.x * (RES_WARPS * 32u); i0 < n_def; i0 += gridDim.x * (RES_WARPS * 32u)) {
        const uint32_t i = i0 + threadIdx.x;
        uint32_t g = 0;
        unsigned long long carry = 0;
        bool alive = false;
        if (i < n_def) {
            g = P.deferred[i];
            // no top-level line between the start of g's range and g: the carry into the range governs
            for (long long r = (long long)(g / RCH) - 1;; r--) {
                const unsigned long long sv = r >= 0 ? P.range_state[r] : (ST_PREFIX | P.carry_in);
                if ((sv & ST_MASK) == ST_PREFIX) {
                    carry = sv & ~ST_MASK;
                    break;
                }
            }
            alive = (carry & CV_HAS_TOP) && (carry & CV_VOK) &&
                    P.tab.vendor_first[(uint32_t)(carry >> 44) & 0xffffu] >= (carry & CV_ANCHOR_MASK);  // vendor_first is final here
        }
        // alive: the warp stages the chunk again; head lines = device lines in front of its first top-level line
        uint32_t am = __ballot_sync(0xffffffffu, alive);
        while (am) {
            const int src = __ffs((int)am) - 1;
            am &= am - 1u;
            const uint32_t gg = __shfl_sync(0xffffffffu, g, src);
            const unsigned long long cc = __shfl_sync(0xffffffffu, carry, src);
            const uint32_t key_hi = ((uint32_t)(cc >> 44) & 0xffffu) << 16;
            const unsigned long long anchor = cc & CV_ANCHOR_MASK;
            const uint32_t n_rel = stage_chunk_manual(P.text, P.n, gg, lane, stg[w]);
            uint32_t kh[2], th[2], rawnl;
            chunk_masks(st, lane, n_rel, k7f, k0a, k80, kh, th, rawnl);
            const uint32_t bal0 = __ballot_sync(0xffffffffu, th[0] != 0u);
            const uint32_t bal1 = __ballot_sync(0xffffffffu, th[1] != 0u);
            const uint32_t pre0 = kh[0] & ~th[0] & (th[0] ? (th[0] & (0u - th[0])) - 1u : 0xffffffffu);
            const uint32_t pre1 = kh[1] & ~th[1] & (th[1] ? (th[1] & (0u - th[1])) - 1u : 0xffffffffu);
            const unsigned long long cbase = P.base + (unsigned long long)gg * CW;
            if ((bal0 & lt_mask) == 0u) fold_lines(P.tab, st, cbase, pre0, lane * 32u + 1u, key_hi, anchor);
            if (bal0 == 0u && (bal1 & lt_mask) == 0u) fold_lines(P.tab, st, cbase, pre1, (uint32_t)HALF + lane * 32u + 1u, key_hi, anchor);
            __syncwarp();
        }
    }
}

}  // namespace kxparse4
