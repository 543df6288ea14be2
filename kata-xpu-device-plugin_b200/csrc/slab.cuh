// slab.cuh -- layout of the per-rank slab exchanged by the one ncclAllGather of the sharded
// pci.ids load (comm.cu packs it, api.cu merges the gathered slabs).
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace kxcomm {

struct SlabHeader {
    uint32_t n_rows, n_vendors, blob_bytes, overflow;  // overflow: a capacity of this slab was exceeded
    unsigned long long trunc;                          // bufio.ErrTooLong cut-off of the shard (all-ones = none)
    unsigned long long reserved;
};
struct SlabRow {     // 32 bytes: one candidate (vendor,device) row of the shard
    uint32_t key, name_len;
    unsigned long long line, anchor;
    uint32_t name_off, pad;
};
struct SlabVendor {  // 16 bytes: earliest top-level line of the shard with this 4-hex prefix
    uint32_t vendor, pad;
    unsigned long long first;
};
struct SlabCaps { uint32_t rows, vendors, blob; };

__host__ __device__ static inline size_t slab_rows_off() { return sizeof(SlabHeader); }
__host__ __device__ static inline size_t slab_vendors_off(const SlabCaps &c) { return sizeof(SlabHeader) + (size_t)c.rows * sizeof(SlabRow); }
__host__ __device__ static inline size_t slab_blob_off(const SlabCaps &c) { return slab_vendors_off(c) + (size_t)c.vendors * sizeof(SlabVendor); }
__host__ __device__ static inline size_t slab_bytes(const SlabCaps &c) { return slab_blob_off(c) + c.blob; }

}  // namespace kxcomm
