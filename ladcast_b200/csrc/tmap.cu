// Host-side TMA tensor-map construction.  cuTensorMapEncodeTiled is fetched through the runtime
// (cudaGetDriverEntryPoint) so the library has no link-time dependency on libcuda.
#include "tmap.h"

#include <cstring>
#include <mutex>
#include <unordered_map>

namespace lc {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct Key {
  const void* ptr;
  uint64_t dims[4];
  uint64_t strides[3];
  uint32_t box[4];
  int rank;
  bool operator==(const Key& o) const { return std::memcmp(this, &o, sizeof(Key)) == 0; }
};
struct KeyHash {
  size_t operator()(const Key& k) const {
    const uint64_t* p = reinterpret_cast<const uint64_t*>(&k);
    size_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(Key) / 8; ++i) h = (h ^ p[i]) * 1099511628211ull;
    return h;
  }
};

static std::mutex g_mu;
static std::unordered_map<Key, CUtensorMap, KeyHash> g_cache;

int make_tmap_bf16(CUtensorMap* out, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box) {
  LC_REQUIRE(rank >= 2 && rank <= 4, "tensor map rank must be 2..4");
  LC_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA base pointer must be 16-byte aligned");
  Key k;
  std::memset(&k, 0, sizeof(k));
  k.ptr = ptr;
  k.rank = rank;
  for (int i = 0; i < rank; ++i) {
    k.dims[i] = dims[i];
    k.box[i] = box[i];
    if (i < rank - 1) {
      k.strides[i] = strides_bytes[i];
      LC_REQUIRE(strides_bytes[i] % 16 == 0, "TMA global strides must be multiples of 16 bytes");
    }
  }
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_cache.find(k);
    if (it != g_cache.end()) {
      *out = it->second;
      return 0;
    }
  }
  EncodeTiledFn fn = get_encode_fn();
  LC_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  cuuint64_t gd[4];
  cuuint64_t gs[3];
  cuuint32_t bx[4], es[4];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i < rank - 1) gs[i] = strides_bytes[i];
  }
  CUtensorMap m;
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(ptr), gd, gs,
                  bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r)) + " (rank " +
              std::to_string(rank) + ", dims " + std::to_string(dims[0]) + "x" + std::to_string(dims[1]) + ", box " +
              std::to_string(box[0]) + "x" + std::to_string(box[1]) + ")");
    return -3;
  }
  {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_cache.size() > 65536) g_cache.clear();
    g_cache.emplace(k, m);
  }
  *out = m;
  return 0;
}

int make_tmap_2d_bf16(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                      uint32_t box_inner, uint32_t box_outer) {
  uint64_t dims[2] = {inner, outer};
  uint64_t strides[1] = {pitch_bytes};
  uint32_t box[2] = {box_inner, box_outer};
  return make_tmap_bf16(out, ptr, 2, dims, strides, box);
}

void tmap_cache_clear() {
  std::lock_guard<std::mutex> lk(g_mu);
  g_cache.clear();
}

}  // namespace lc
