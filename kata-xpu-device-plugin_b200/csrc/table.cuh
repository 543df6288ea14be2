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

// counters[] indices.  Four 128-byte lines: words that are hammered with atomics at the same time sit on
// different lines, and the flags that are only polled (OVERFLOW: once per range / task by every warp) have
// a line of their own -- a load of a line under same-address atomic fire queues behind the atomics (the
// poll was 36 % of the stall samples of resolve_chunks_kernel when it shared a line with NKEYS).
#define KX_C_TICKET 0     // line 0: parse range tickets
#define KX_C_DEFER 1      //         resolve queue length
#define KX_C_GRIDBAR 2    //         small-text kernel: grid barrier arrivals
#define KX_C_NKEYS 32     // line 1: claimed slots (parse, resolve, merge)
#define KX_C_NSEL 33      //         row handles (finalize)
#define KX_C_BLOB_CURSOR 64   // line 2: name bytes (finalize)
#define KX_C_OVERFLOW 96      // line 3: flags and results, written rarely
#define KX_C_LONGLINE_HINT 97
#define KX_C_NROWS 98
#define KX_C_BLOB_OVERFLOW 99
#define KX_C_NEED_TRUNC 100
#define KX_C_XSTATUS 101   // sharded load: KX_XS_* bits, identical on every rank after the exchange
#define KX_C_XROWS 102     // sharded load: winner rows of all ranks
#define KX_C_XBLOB 103     // sharded load: name bytes of all ranks
#define KX_C_XMAXKEYS 104  // sharded load: max over ranks of the local candidate key count
#define KX_C_COUNT 128

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
