// common.cuh -- context, error plumbing and small device helpers shared by the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <mutex>

#include "../../include/kxpu.h"

struct kxpu_ctx {
    int device = -1;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[2 * KXPU_T_COUNT] = {};
    cudaEvent_t ev_user[2] = {};
    bool ev_used[KXPU_T_COUNT] = {};
    int force_rch = 0;         // KXPU_RCH: chunks per parse range (0 = automatic), for tests
    bool stage_timing = true;  // per-stage CUDA events (kxpu_last_timings); kxpu_set_stage_timing(ctx, 0) drops them
    uint64_t launches = 0;
    std::mutex mu;
    char err[512] = {0};
    // pinned staging for small H2D/D2H control words
    uint32_t *h_ctl = nullptr;  // 64 words, pinned
    // NCCL (lazy)
    void *nccl_comm = nullptr;
    int nranks = 1, rank = 0;
    // peer-memory exchange of the sharded load (comm.cu): this rank's region and the peers' regions
    // mapped through CUDA IPC; layout [flags u32[2][16] | pad to 256 | slab[2][nranks]]
    static constexpr int KX_P2P_MAX_RANKS = 16;
    uint8_t *p2p_local = nullptr;
    uint8_t *p2p_peer[KX_P2P_MAX_RANKS] = {};
    size_t p2p_stride = 0;
    uint32_t p2p_epoch = 0;
    uint32_t *p2p_scratch = nullptr;  // device: [0] push-done counter, [1] wait timeout flag
    bool p2p_ok = false;
    int parse_version = 2;
};

#define KX_SET_ERR(ctx, ...) snprintf((ctx)->err, sizeof((ctx)->err), __VA_ARGS__)

#define KX_CUDA(ctx, call)                                                                  \
    do {                                                                                    \
        cudaError_t e__ = (call);                                                           \
        if (e__ != cudaSuccess) {                                                           \
            KX_SET_ERR(ctx, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return KXPU_E_CUDA;                                                             \
        }                                                                                   \
    } while (0)

#define KX_LAUNCHED(ctx) ((ctx)->launches++)

struct KxTimer {  // records a CUDA-event pair around a stage on the ctx stream
    kxpu_ctx *c;
    int idx;
    bool open = true;
    KxTimer(kxpu_ctx *ctx, int i) : c(ctx), idx(i) {
        open = c->stage_timing;
        if (open) cudaEventRecord(c->ev[2 * i], c->stream);
    }
    void stop() {
        if (open) { cudaEventRecord(c->ev[2 * idx + 1], c->stream); c->ev_used[idx] = true; open = false; }
    }
    ~KxTimer() { stop(); }
};

static inline void kx_clear_timings(kxpu_ctx *c) { memset(c->ev_used, 0, sizeof(c->ev_used)); }

// ------------------------------------------------------------------ device helpers
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t kx_lane() { return threadIdx.x & 31u; }

// Parse four ASCII bytes (first character in the low byte) as lowercase hex.
// Only [0-9a-f] is accepted: sysfs ids are lowercase, and the reference compares raw
// bytes (strings.HasPrefix, device_plugin.go:237,265).
__device__ __forceinline__ bool kx_hex4(uint32_t w, uint32_t &v) {
    uint32_t acc = 0, bad = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint32_t c = (w >> (8 * i)) & 0xffu;
        uint32_t d = c - 0x30u, a = c - 0x61u;
        uint32_t x = d < 10u ? d : a + 10u;
        bad |= (d >= 10u) & (a >= 6u);
        acc = (acc << 4) | (x & 15u);
    }
    v = acc;
    return bad == 0;
}

__device__ __forceinline__ uint32_t kx_hash(uint32_t key) { return key * 0x9E3779B1u; }
#endif
