// table.cuh -- device layout of the (vendor,device) table built from a pci.ids text.
//
// HBM layout (one arena, see kxpu_table in pciids_api.cu):
//   keys        u32 [cap+1]   (vendor<<16)|device, 0xffffffff = empty; slot `cap` is the
//                             dedicated slot of key 0xffffffff itself
//   min_line    u64 [cap+1]   smallest global offset of a "\t"+device line seen under a
//                             top-level line with prefix vendor
//   min_anchor  u64 [cap+1]   smallest global offset of such a governing top-level line
//   vendor_first u64 [65536]  smallest global offset of a top-level line with that prefix
//   row_of_slot i32 [cap+1]   row handle after finalize, -1 = not a valid hit
// A slot is a valid hit iff min_anchor == vendor_first[vendor] (the device line sits in
// the block of the FIRST matching vendor line, device_plugin.go:263-267) and
// min_line < trunc (bufio.ErrTooLong cut-off).  Because vendor blocks are disjoint and
// ordered, min_line and min_anchor always stem from the same block.
#pragma once
#include "common.cuh"

#define KX_EMPTY_KEY 0xFFFFFFFFu
#define KX_NO_OFF 0xFFFFFFFFFFFFFFFFull

// counters[] indices
#define KX_C_TICKET 0
#define KX_C_NKEYS 1
#define KX_C_OVERFLOW 2
#define KX_C_LONGLINE_HINT 3
#define KX_C_NROWS 4
#define KX_C_BLOB_CURSOR 5
#define KX_C_BLOB_OVERFLOW 6
#define KX_C_NEED_TRUNC 7
#define KX_C_PEND_OVERFLOW 8
#define KX_C_NSEL 9
#define KX_C_DEFER 10
#define KX_C_SLAB_OVERFLOW 11
#define KX_C_COUNT 16

struct KxTableDev {
    uint32_t *keys;
    unsigned long long *min_line;
    unsigned long long *min_anchor;
    unsigned long long *vendor_first;
    uint32_t *counters;
    unsigned long long *trunc;  // [1] global offset where the reference's scan stops (KX_NO_OFF = never)
    uint32_t cap;               // power of two
    uint32_t shift;             // 32 - log2(cap)
    uint32_t max_keys;          // growth threshold
};
