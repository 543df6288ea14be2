// table.cuh -- device layout of the (vendor,device) table built from a pci.ids text.
//
// HBM layout (one arena per table, pooled per ctx, see api.cu):
//   slots        KxSlot [cap+1]  open addressing, ONE 32-byte sector per slot, so a fold, a probe
//                                or a merge insert touches a single sector:
//                  key         (vendor<<16)|device, 0xffffffff = empty; slot `cap` is the
//                              dedicated slot of key 0xffffffff itself
//                  row         row handle after finalize / merge, -1 = not a valid hit
//                  min_line    smallest global offset of a "\t"+device line seen under a
//                              top-level line with prefix vendor
//                  min_anchor  smallest global offset of such a governing top-level line
//   vendor_first u64 [65536]   smallest global offset of a top-level line with that prefix
//   trunc        u64           bufio.ErrTooLong cut-off (all-ones = none)
//   counters     u32 [KX_C_COUNT]
// A slot is a valid hit iff min_anchor == vendor_first[vendor] (the device line sits in
// the block of the FIRST matching vendor line, device_plugin.go:263-267) and
// min_line < trunc (bufio.ErrTooLong cut-off).  Because vendor blocks are disjoint and
// ordered, min_line and min_anchor always stem from the same block.
// The whole [slots | vendor_first | trunc] region resets to 0xff bytes, the counters to 0.
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
#define KX_C_GRIDBAR 8   // small-text kernel: grid barrier arrivals
#define KX_C_NSEL 9
#define KX_C_DEFER 10
#define KX_C_XSTATUS 11   // sharded load: KX_XS_* bits, identical on every rank after the exchange
#define KX_C_XROWS 12     // sharded load: winner rows of all ranks
#define KX_C_XBLOB 13     // sharded load: name bytes of all ranks
#define KX_C_XMAXKEYS 14  // sharded load: max over ranks of the local candidate key count
#define KX_C_COUNT 16

struct __align__(32) KxSlot {
    uint32_t key;
    int32_t row;
    unsigned long long min_line;
    unsigned long long min_anchor;
    unsigned long long spare;
};

struct KxTableDev {
    KxSlot *slots;
    unsigned long long *vendor_first;
    uint32_t *counters;
    unsigned long long *trunc;  // [1] global offset where the reference's scan stops (KX_NO_OFF = never)
    uint32_t cap;               // power of two
    uint32_t shift;             // 32 - log2(cap)
    uint32_t max_keys;          // growth threshold
};
