// api_pending.cu — entry points whose device kernels are not built yet: they fail loudly (no CPU fallback).
#include "../../include/swcgpu.h"

#define PENDING(...) { return SWC_ERR_UNSUPPORTED; }
extern "C" {
int32_t swc_lzma_decompress(const uint8_t *, size_t, uint8_t **, size_t *, size_t *) PENDING()
int32_t swc_lzma_decompress_raw(const uint8_t *, size_t, int32_t, int32_t, int32_t, int64_t, int64_t, uint8_t **, size_t *, size_t *) PENDING()
int32_t swc_lzma2_decompress(const uint8_t *, size_t, uint8_t **, size_t *, size_t *) PENDING()
int32_t swc_lzma2_decompress_batch(const uint8_t *, const uint64_t *, const uint64_t *, const uint8_t *, uint8_t *, const uint64_t *,
                                   const uint64_t *, uint64_t *, uint64_t *, int32_t *, uint64_t, void *) PENDING()
int32_t swc_xz_unarchive(const uint8_t *, size_t, uint8_t **, size_t *) PENDING()
int32_t swc_xz_split_unarchive(const uint8_t *, size_t, uint8_t **, size_t *, size_t **, size_t *) PENDING()
}
