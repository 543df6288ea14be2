#include "common.cuh"
extern "C" int32_t kxpu_comm_unique_id(uint8_t *) { return KXPU_E_NCCL; }
extern "C" int32_t kxpu_comm_init(kxpu_ctx *, int32_t, int32_t, const uint8_t *) { return KXPU_E_NCCL; }
extern "C" int32_t kxpu_comm_destroy(kxpu_ctx *) { return KXPU_E_NCCL; }
extern "C" int32_t kxpu_pciids_load_sharded(kxpu_ctx *, const void *, size_t, uint64_t, kxpu_table **) { return KXPU_E_NCCL; }
