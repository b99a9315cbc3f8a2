// TMA tensor-map helpers (host).  All maps are bf16, 128-byte swizzle, zero OOB fill.
#pragma once
#include "common.cuh"

namespace lc {
int make_tmap_bf16(CUtensorMap* out, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box);
int make_tmap_2d_bf16(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                      uint32_t box_inner, uint32_t box_outer);
void tmap_cache_clear();
}  // namespace lc
