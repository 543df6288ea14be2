// finalize.cuh -- the kernels behind the parse: bufio.ErrTooLong cut-off, validity + name
// sanitiser (K4, device_plugin.go:241-251), batched join (K3) and name gather.
#pragma once
#include "exchange.cuh"
#include "parse_common.cuh"

namespace kxparse {

// ------------------------------------------------------------------------------
// bufio.ErrTooLong cut-off (slow path, only when the parse kernel raised the hint).
// trunc = global offset of the first line whose content is >= 65536 bytes.
// One CTA; every thread scans a contiguous byte range for newlines and reports the
// first/last newline and the longest gap inside; thread 0 stitches the ranges.
// ------------------------------------------------------------------------------
constexpr unsigned long long MAX_TOKEN = 65536ull;

__global__ void __launch_bounds__(1024) trunc_kernel(const uint8_t *__restrict__ text, unsigned long long n,
                                                      unsigned long long base, unsigned long long *trunc_out, uint32_t *counters) {
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
        counters[KX_C_NEED_TRUNC] = 1u;  // the cut-off is there: finalize_select may proceed
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
    // validity is judged against these (single text: the table's own; sharded load: the
    // all-reduced minima of every rank, comm.cu)
    const unsigned long long *vendor_first;
    const unsigned long long *trunc;
    uint32_t *row_key;
    unsigned long long *row_line;
    unsigned long long *row_anchor;
    uint32_t *row_name_off;
    uint32_t *row_name_len;
    uint32_t *sel;  // [cap+1] valid slots (stage 1 -> stage 2)
    uint8_t *blob;
    uint32_t blob_cap;
    kxx::WaitSpec wait;  // sharded load: the all-reduced minima are complete when these flags are up
};

// Stage 1, one thread per table slot: validity; valid slots are compacted into F.sel.
// KX_C_NEED_TRUNC: 0 = cut-off never computed, 1 = computed (host ran trunc_kernel), 2 = asked
// for.  When the parse raised the long-line hint and the cut-off is not there yet, every block
// leaves (the test does not depend on what block 0 writes) and the host finalizes again.
__global__ void __launch_bounds__(256) finalize_select_kernel(const FinalizeParams F) {
    kxx::wait_flags_cta(F.wait);
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (F.tab.counters[KX_C_LONGLINE_HINT] != 0u && F.tab.counters[KX_C_NEED_TRUNC] != 1u) {
        if (slot == 0) F.tab.counters[KX_C_NEED_TRUNC] = 2u;
        return;
    }
    bool valid = false;
    if (slot <= F.tab.cap) {
        const uint4 head = *reinterpret_cast<const uint4 *>(&F.tab.slots[slot]);
        const unsigned long long line = ((unsigned long long)head.w << 32) | head.z;
        const uint32_t key = slot == F.tab.cap ? KX_EMPTY_KEY : head.x;
        valid = line != KX_NO_OFF && !(slot < F.tab.cap && key == KX_EMPTY_KEY);
        if (valid) valid = F.tab.slots[slot].min_anchor == F.vendor_first[key >> 16] && line < *F.trunc;
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
// blob.  Persistent grid: a CTA takes groups of 8 selected slots until the (device-side) count is
// used up, so a load that selects nothing costs one wave of CTAs that read the count and leave.
// Blob space is claimed once per group (8 names) to keep the cursor atomic cheap.
__global__ void __launch_bounds__(FIN_WARPS * 32) finalize_kernel(const FinalizeParams F) {
    __shared__ uint8_t s_buf[FIN_WARPS][NAME_BUF + 32];
    __shared__ uint32_t s_len[FIN_WARPS];
    __shared__ uint32_t s_base;
    const uint32_t lane = threadIdx.x & 31u, wl = threadIdx.x >> 5;
    const uint32_t nsel = F.tab.counters[KX_C_NSEL];
    uint8_t *buf = s_buf[wl];
    for (uint32_t g0 = blockIdx.x * FIN_WARPS; g0 < nsel; g0 += gridDim.x * FIN_WARPS) {
        const uint32_t si = g0 + wl;
        const bool active = si < nsel;
        uint32_t slot = 0, key = 0, len = 0, start = 0, end = 0, out_len = 0;
        unsigned long long line = 0, anchor = 0, rs = 0;
        bool fast = false;
        if (active) {
            slot = F.sel[si];
            line = F.tab.slots[slot].min_line;
            key = slot == F.tab.cap ? KX_EMPTY_KEY : F.tab.slots[slot].key;
            anchor = F.tab.slots[slot].min_anchor;
            rs = line - F.base + 5ull;  // rest of the line after "\t" + 4 hex digits
            bool found = false;
            for (uint32_t o = 0; o < (uint32_t)NAME_BUF + 32u && !found; o += 128u) {
                // 128 bytes per step: most names end inside the first one
                uint32_t nlm[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const unsigned long long pos = rs + o + 32u * k + lane;
                    const uint32_t c = pos < F.n ? F.text[pos] : 0x0au;  // EOF terminates the last line
                    if (o + 32u * k < (uint32_t)NAME_BUF + 32u) buf[o + 32u * k + lane] = (uint8_t)c;
                    nlm[k] = __ballot_sync(0xffffffffu, c == 0x0au);
                }
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (!found && nlm[k]) { len = o + 32u * k + (uint32_t)__ffs((int)nlm[k]) - 1u; found = true; }
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
        if (active) {
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
                F.tab.slots[slot].row = (int32_t)si;
                F.row_key[si] = key;
                F.row_line[si] = line;
                F.row_anchor[si] = anchor;
                F.row_name_off[si] = room ? out_off : 0u;
                F.row_name_len[si] = room ? out_len : 0u;
            }
        }
        __syncthreads();  // s_len / s_base / the staging buffers are reused by the next group
    }
}

// ------------------------------------------------------------------------------
// K3 batched join: one thread per key, probe the table (L2 resident; key and row handle share
// one 32-byte sector), return the row handle.
// ------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) lookup_kernel(const uint32_t *__restrict__ keys, size_t n, const KxSlot *__restrict__ slots,
                                                      uint32_t cap, uint32_t shift, int32_t *__restrict__ rows_out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) rows_out[i] = table_probe(slots, cap, shift, keys[i]);
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

// [slots | vendor_first | trunc] <- 0xff, counters <- 0: one launch per table arena (api.cu)
__global__ void __launch_bounds__(256) arena_reset_kernel(uint4 *ff, size_t n_ff16, uint32_t *counters) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t k = i; k < n_ff16; k += stride) ff[k] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    if (i < KX_C_COUNT) counters[i] = 0u;
}

}  // namespace kxparse
