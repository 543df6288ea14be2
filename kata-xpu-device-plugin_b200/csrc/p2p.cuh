// p2p.cuh -- peer-memory exchange of the sharded pci.ids load (included by comm.cu).
//
// The fused "pack + all-gather": every rank maps the exchange region of every other rank (CUDA IPC,
// once, inside kxpu_comm_init).  Per load a rank packs its slab into its own slot, PUSHES it into the
// same slot of every peer over NVLink (plain 16-byte stores into peer memory), then raises its flag in
// every peer's flag array; the merge waits until all flags of this load have arrived.  No NCCL call,
// no staging copy.  Two buffers alternate by load number: a rank can only start load e + 2 after every
// peer delivered load e + 1, i.e. after every peer finished merging load e.
// Region layout: [flags u32[2][16] | pad to 256 B | slab[2][nranks]].
#pragma once
#include "common.cuh"
#include "slab.cuh"
#include "table.cuh"

namespace kxcomm {

constexpr size_t P2P_FLAGS_BYTES = 256;
static const SlabCaps kDefaultCaps{32768u, 8192u, 1u << 20};

struct PushParams {
    uint8_t *dst[kxpu_ctx::KX_P2P_MAX_RANKS];    // slot of this rank in every rank's region (own region included)
    uint32_t *flag[kxpu_ctx::KX_P2P_MAX_RANKS];  // my flag word of this buffer in every rank's region
    int nranks;
    uint32_t epoch;
    SlabCaps caps;
    uint32_t *scratch;  // device: [0] CTAs done, [2] vendor rows appended, [3] overflow bits
    // the local table (kx_table_local_view)
    uint32_t n_rows, blob_used;
    const uint32_t *row_key, *row_name_off, *row_name_len;
    const unsigned long long *row_line, *row_anchor, *trunc, *vendor_first;
    const uint8_t *blob;
};

// pack + push in one kernel: candidate rows, vendor rows and names go straight from the local table
// into this rank's slot of EVERY rank's exchange region; the last CTA writes the header and raises
// the flags.
__global__ void __launch_bounds__(256) pack_push_kernel(const PushParams P) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    const uint32_t nr = P.n_rows < P.caps.rows ? P.n_rows : P.caps.rows;
    for (size_t i = tid; i < nr; i += nth) {
        SlabRow r;
        r.key = P.row_key[i]; r.name_len = P.row_name_len[i]; r.line = P.row_line[i]; r.anchor = P.row_anchor[i];
        r.name_off = P.row_name_off[i]; r.pad = 0;
        for (int q = 0; q < P.nranks; q++) reinterpret_cast<SlabRow *>(P.dst[q] + slab_rows_off())[i] = r;
    }
    for (size_t v = tid; v < 65536; v += nth) {
        const unsigned long long f = P.vendor_first[v];
        if (f == KX_NO_OFF) continue;
        const uint32_t idx = atomicAdd(&P.scratch[2], 1u);
        if (idx >= P.caps.vendors) { atomicOr(&P.scratch[3], 2u); continue; }
        SlabVendor sv;
        sv.vendor = (uint32_t)v; sv.pad = 0; sv.first = f;
        for (int q = 0; q < P.nranks; q++) reinterpret_cast<SlabVendor *>(P.dst[q] + slab_vendors_off(P.caps))[idx] = sv;
    }
    if (P.blob_used <= P.caps.blob) {
        const size_t n16 = ((size_t)P.blob_used + 15) / 16;  // the blob lives in a 256-byte aligned arena with slack
        const uint4 *s4 = reinterpret_cast<const uint4 *>(P.blob);
        for (size_t i = tid; i < n16; i += nth) {
            const uint4 x = s4[i];
            for (int q = 0; q < P.nranks; q++) reinterpret_cast<uint4 *>(P.dst[q] + slab_blob_off(P.caps))[i] = x;
        }
    }
    // last CTA: header, then my flag everywhere
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t prev = atomicAdd(&P.scratch[0], 1u);
        if (prev == gridDim.x - 1u) {
            __threadfence_system();
            SlabHeader h;
            const uint32_t nv = *reinterpret_cast<volatile uint32_t *>(&P.scratch[2]);
            h.n_rows = nr;
            h.n_vendors = nv < P.caps.vendors ? nv : P.caps.vendors;
            h.blob_bytes = P.blob_used;
            h.overflow = *reinterpret_cast<volatile uint32_t *>(&P.scratch[3]) |
                         ((P.n_rows > P.caps.rows || P.blob_used > P.caps.blob) ? 1u : 0u);
            h.trunc = *P.trunc;
            h.reserved = 0;
            for (int q = 0; q < P.nranks; q++) *reinterpret_cast<SlabHeader *>(P.dst[q]) = h;
            P.scratch[0] = 0u; P.scratch[2] = 0u; P.scratch[3] = 0u;
            __threadfence_system();
            for (int q = 0; q < P.nranks; q++) *reinterpret_cast<volatile uint32_t *>(P.flag[q]) = P.epoch;
            __threadfence_system();
        }
    }
}

// one warp: lane q waits for rank q's flag of this load
__global__ void wait_flags_kernel(const uint32_t *flags, int nranks, uint32_t epoch, uint32_t *timeout_flag) {
    const int q = threadIdx.x;
    if (q < nranks) {
        const long long t0 = clock64();
        while (*reinterpret_cast<const volatile uint32_t *>(&flags[q]) < epoch) {
            __nanosleep(200);
            if (clock64() - t0 > 8000000000ll) { *timeout_flag = 1u; break; }  // ~4 s: a peer died
        }
    }
    __threadfence_system();
}

// Map every peer's exchange region.  Any failure leaves p2p_ok false on EVERY rank (the ranks agree
// through one more tiny all-gather) and the ncclAllGather path is used.
template <typename AllGather>
static void p2p_setup(kxpu_ctx *ctx, AllGather all_gather) {
    ctx->p2p_ok = false;
    const int R = ctx->nranks;
    if (R < 2 || R > kxpu_ctx::KX_P2P_MAX_RANKS || getenv("KXPU_NO_P2P")) return;
    const size_t stride = (slab_bytes(kDefaultCaps) + 255) / 256 * 256;
    const size_t bytes = P2P_FLAGS_BYTES + 2 * (size_t)R * stride;
    bool ok = true;
    uint8_t *local = nullptr, *d_x = nullptr;
    cudaIpcMemHandle_t mine, all[kxpu_ctx::KX_P2P_MAX_RANKS];
    memset(&mine, 0, sizeof mine);
    ok = ok && cudaMalloc((void **)&local, bytes) == cudaSuccess;
    ok = ok && cudaMemset(local, 0, P2P_FLAGS_BYTES) == cudaSuccess;
    ok = ok && cudaMalloc((void **)&ctx->p2p_scratch, 256) == cudaSuccess && cudaMemset(ctx->p2p_scratch, 0, 256) == cudaSuccess;
    ok = ok && cudaIpcGetMemHandle(&mine, local) == cudaSuccess;
    // exchange the handles with the communicator that exists already
    const size_t hb = sizeof(cudaIpcMemHandle_t);
    if (cudaMalloc((void **)&d_x, hb * (size_t)(R + 1) + 64) != cudaSuccess) { d_x = nullptr; ok = false; }
    uint8_t okbyte[kxpu_ctx::KX_P2P_MAX_RANKS + 1] = {};
    if (d_x) {
        cudaMemcpyAsync(d_x, &mine, hb, cudaMemcpyHostToDevice, ctx->stream);
        const int nrc = all_gather(d_x, d_x + hb, hb);
        cudaMemcpyAsync(all, d_x + hb, hb * (size_t)R, cudaMemcpyDeviceToHost, ctx->stream);
        if (nrc != 0 || cudaStreamSynchronize(ctx->stream) != cudaSuccess) ok = false;
    }
    if (ok) {
        for (int q = 0; q < R && ok; q++) {
            if (q == ctx->rank) { ctx->p2p_peer[q] = local; continue; }
            void *pp = nullptr;
            if (cudaIpcOpenMemHandle(&pp, all[q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; cudaGetLastError(); break; }
            ctx->p2p_peer[q] = (uint8_t *)pp;
        }
    }
    if (d_x) {  // agreement: all ranks or none
        const uint8_t mineok = ok ? 1 : 0;
        uint8_t *d_ok = d_x + hb * (size_t)(R + 1);
        cudaMemcpyAsync(d_ok, &mineok, 1, cudaMemcpyHostToDevice, ctx->stream);
        const int nrc = all_gather(d_ok, d_ok + 16, 1);
        cudaMemcpyAsync(okbyte, d_ok + 16, (size_t)R, cudaMemcpyDeviceToHost, ctx->stream);
        if (nrc != 0 || cudaStreamSynchronize(ctx->stream) != cudaSuccess) ok = false;
        for (int q = 0; q < R; q++) ok = ok && okbyte[q] == 1;
        cudaFree(d_x);
    }
    if (!ok) {
        for (int q = 0; q < R; q++)
            if (q != ctx->rank && ctx->p2p_peer[q]) cudaIpcCloseMemHandle(ctx->p2p_peer[q]);
        memset(ctx->p2p_peer, 0, sizeof ctx->p2p_peer);
        if (local) cudaFree(local);
        if (ctx->p2p_scratch) { cudaFree(ctx->p2p_scratch); ctx->p2p_scratch = nullptr; }
        cudaGetLastError();
        if (getenv("KXPU_TRACE_MERGE")) fprintf(stderr, "[kxpu] rank %d: peer-memory exchange unavailable, using ncclAllGather\n", ctx->rank);
        return;
    }
    ctx->p2p_local = local;
    ctx->p2p_stride = stride;
    ctx->p2p_epoch = 0;
    ctx->p2p_ok = true;
    if (getenv("KXPU_TRACE_MERGE")) fprintf(stderr, "[kxpu] rank %d: peer-memory exchange over %d ranks, %zu B per rank\n", ctx->rank, R, bytes);
}

static void p2p_teardown(kxpu_ctx *ctx) {
    if (!ctx->p2p_local) return;
    for (int q = 0; q < ctx->nranks; q++)
        if (q != ctx->rank && ctx->p2p_peer[q]) cudaIpcCloseMemHandle(ctx->p2p_peer[q]);
    memset(ctx->p2p_peer, 0, sizeof ctx->p2p_peer);
    cudaFree(ctx->p2p_local);
    ctx->p2p_local = nullptr;
    if (ctx->p2p_scratch) { cudaFree(ctx->p2p_scratch); ctx->p2p_scratch = nullptr; }
    ctx->p2p_ok = false;
}

}  // namespace kxcomm
