#include "common.cuh"
extern "C" int32_t kxpu_cdi_emit(kxpu_ctx *, int32_t, const kxpu_cdidev *, size_t, uint8_t *, size_t, size_t *) { return KXPU_E_UNSUPPORTED; }
extern "C" int32_t kxpu_alloc_names(kxpu_ctx *, const uint64_t *, size_t, uint8_t *, size_t, uint32_t *, size_t *) { return KXPU_E_UNSUPPORTED; }
extern "C" int32_t kxpu_lw_encode(kxpu_ctx *, const uint32_t *, const uint8_t *, size_t, uint8_t *, size_t, size_t *) { return KXPU_E_UNSUPPORTED; }
