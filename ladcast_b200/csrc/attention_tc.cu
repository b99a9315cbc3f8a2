// tcgen05 flash attention, head_dim 128, non-causal, no mask (F.scaled_dot_product_attention at
// LaDCast_3D_model.py:199-201).  Q/K/V are read by TMA straight out of the token-major [B*S, 3d] projection
// buffer (q | k | v), so no head-major transposes exist anywhere.
//
//   CTA = one (sample, head, 256-query block) = two 128-query tiles A and B that share every K/V tile; 320 threads:
//     warp 0    : TMA producer — Q (both tiles) once, K_j / V_j tiles of 64 keys through a 4-stage ring, so loads run
//                 ~3 tiles ahead of the tensor pipe (L2->smem latency under load is ~2 us, longer than one tile)
//     warp 1    : TMEM allocator + single-thread tcgen05.mma issuer.  Per KV tile: S_A = Q_A K^T, S_B = Q_B K^T
//                 (double-buffered in TMEM, issued one tile ahead), O_A += P_A V, O_B += P_B V (V consumed MN-major,
//                 i.e. exactly as it lies in memory)
//     warps 2-5 : softmax of tile A, warps 6-9: softmax of tile B.  One thread per query row reads its S row from
//                 TMEM (no shuffles), online softmax in the log2 domain (one FFMA + one MUFU.EX2 per element), lazy
//                 rescale of O (only when the row max grows by more than 2^8), P (bf16) into a 128B-swizzled smem tile
//                 that is the A operand of the PV product.  While one tile's softmax runs, the tensor pipe works on
//                 the other tile.  Final O / l epilogue per tile.
#include "kernels.h"
#include "ptx.cuh"
#include "tmap.h"

namespace lc {
namespace {

constexpr int HD = 128;
constexpr int BQ = 128;                    // rows per query tile; two tiles per CTA
constexpr int BKV = 64;
constexpr int STAGES = 4;
constexpr int Q_BYTES = 128 * 128 * 2;     // 32 KB per query tile: two [128 rows][64 dims] swizzled boxes
constexpr int QSUB_BYTES = 128 * 64 * 2;   // 16 KB
constexpr int KV_BYTES = BKV * 128 * 2;    // 16 KB: two [64 keys][64 dims] swizzled boxes
constexpr int KVSUB_BYTES = BKV * 64 * 2;  // 8 KB
constexpr int P_BYTES = 128 * BKV * 2;     // 16 KB: [128 rows][64 keys]
constexpr int SMEM_BYTES = 2 * Q_BYTES + 2 * P_BYTES + 2 * STAGES * KV_BYTES + 1024 + 256;
constexpr int NUM_THREADS = 320;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t COL_S = 0;    // S[tile][buf] at COL_S + tile*128 + buf*64
constexpr uint32_t COL_O = 256;  // O[tile] at COL_O + tile*128
constexpr float RESCALE_THRESHOLD = 8.0f;

__global__ void __launch_bounds__(NUM_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmkv, int S, int heads,
                    bf16* __restrict__ out_p, int Np, bf16* __restrict__ out_c) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sQ = smem;                       // [2][Q_BYTES]
  uint8_t* sP = sQ + 2 * Q_BYTES;           // [2][P_BYTES]
  uint8_t* sK = sP + 2 * P_BYTES;           // [STAGES][KV_BYTES]
  uint8_t* sV = sK + STAGES * KV_BYTES;     // [STAGES][KV_BYTES]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + STAGES * KV_BYTES);
  uint64_t* q_full = bars;                  // 1
  uint64_t* k_full = bars + 1;              // STAGES
  uint64_t* v_full = k_full + STAGES;       // STAGES
  uint64_t* kv_empty = v_full + STAGES;     // STAGES
  uint64_t* s_full = kv_empty + STAGES;     // [tile][buf] = 4
  uint64_t* p_full = s_full + 4;            // [tile] (4 arrivals: one per softmax warp)
  uint64_t* p_empty = p_full + 2;           // [tile] (PV of the tile's previous KV tile complete)
  uint64_t* o_full = p_empty + 2;           // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d = heads * HD;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * (2 * BQ);
  const int n_tiles = (S + BKV - 1) / BKV;
  const int row_base = b * S;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmq);
    ptx::prefetch_tmap(&tmkv);
    ptx::mbar_init(q_full, 1);
    for (int i = 0; i < STAGES; ++i) {
      ptx::mbar_init(&k_full[i], 1);
      ptx::mbar_init(&v_full[i], 1);
      ptx::mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 4; ++i) ptx::mbar_init(&s_full[i], 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&p_full[i], 4);
      ptx::mbar_init(&p_empty[i], 1);
    }
    ptx::mbar_init(o_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<TMEM_COLS>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    ptx::mbar_expect_tx(q_full, 2 * Q_BYTES);
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      ptx::tma_load_2d(sQ + t * Q_BYTES, &tmq, q_full, h * HD, row_base + q0 + t * BQ);
      ptx::tma_load_2d(sQ + t * Q_BYTES + QSUB_BYTES, &tmq, q_full, h * HD + 64, row_base + q0 + t * BQ);
    }
    int st = 0;
    uint32_t ph = 0;
    for (int j = 0; j < n_tiles; ++j) {
      ptx::mbar_wait(&kv_empty[st], ph ^ 1);
      const int kr = row_base + j * BKV;
      ptx::mbar_expect_tx(&k_full[st], KV_BYTES);
      ptx::tma_load_2d(sK + st * KV_BYTES, &tmkv, &k_full[st], d + h * HD, kr);
      ptx::tma_load_2d(sK + st * KV_BYTES + KVSUB_BYTES, &tmkv, &k_full[st], d + h * HD + 64, kr);
      ptx::mbar_expect_tx(&v_full[st], KV_BYTES);
      ptx::tma_load_2d(sV + st * KV_BYTES, &tmkv, &v_full[st], 2 * d + h * HD, kr);
      ptx::tma_load_2d(sV + st * KV_BYTES + KVSUB_BYTES, &tmkv, &v_full[st], 2 * d + h * HD + 64, kr);
      if (++st == STAGES) { st = 0; ph ^= 1; }
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_qk = ptx::make_idesc_bf16(128, BKV, 0, 0);  // A = Q (K-major), B = K (K-major)
    constexpr uint32_t idesc_pv = ptx::make_idesc_bf16(128, 128, 0, 1);  // A = P (K-major), B = V (MN-major)
    const uint32_t q_addr = ptx::smem_u32(sQ), p_addr = ptx::smem_u32(sP);
    auto issue_qk = [&](int j) {  // both query tiles against K_j
      const int st = j % STAGES;
      ptx::mbar_wait(&k_full[st], (j / STAGES) & 1);
      ptx::tc_fence_after();
      const uint32_t k_addr = ptx::smem_u32(sK + st * KV_BYTES);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const uint32_t d_tmem = tmem_base + COL_S + static_cast<uint32_t>(t * 128 + (j & 1) * BKV);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {  // 128 head dims = 2 sub-tiles x 4 K-steps of 16
          const uint32_t qoff = t * Q_BYTES + (kk >> 2) * QSUB_BYTES + (kk & 3) * 32;
          const uint32_t koff = (kk >> 2) * KVSUB_BYTES + (kk & 3) * 32;
          ptx::umma_f16(d_tmem, ptx::make_smem_desc(q_addr + qoff, 16, 1024),
                        ptx::make_smem_desc(k_addr + koff, 16, 1024), idesc_qk, kk != 0 ? 1u : 0u);
        }
        ptx::umma_commit(&s_full[t * 2 + (j & 1)]);
      }
    };
    ptx::mbar_wait(q_full, 0);
    issue_qk(0);
    for (int j = 0; j < n_tiles; ++j) {
      const int st = j % STAGES;
      if (j + 1 < n_tiles) issue_qk(j + 1);
      ptx::mbar_wait(&v_full[st], (j / STAGES) & 1);
      const uint32_t v_addr = ptx::smem_u32(sV + st * KV_BYTES);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        ptx::mbar_wait(&p_full[t], j & 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < BKV / 16; ++kk) {  // 64 keys = 4 K-steps of 16
          const uint64_t da = ptx::make_smem_desc(p_addr + t * P_BYTES + kk * 32, 16, 1024);
          const uint64_t db = ptx::make_smem_desc(v_addr + kk * 2048, KVSUB_BYTES, 1024);
          ptx::umma_f16(tmem_base + COL_O + static_cast<uint32_t>(t * 128), da, db, idesc_pv, (j | kk) != 0 ? 1u : 0u);
        }
        ptx::umma_commit(&p_empty[t]);
      }
      ptx::umma_commit(&kv_empty[st]);  // K_j and V_j are dead once both tiles' PV products have completed
    }
    ptx::umma_commit(o_full);
  } else if (warp >= 2) {
    // ===================== softmax + epilogue (thread = query row of tile t) =====================
    const int t = (warp - 2) >> 2;  // 0: tile A (warps 2-5), 1: tile B (warps 6-9)
    const int quarter = warp & 3;   // TMEM lane quarter this warp may access (= warp id % 4)
    const int r = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const float scale_log2 = 0.08838834764831845f * 1.4426950408889634f;
    const uint32_t o_addr = tmem_base + lane_addr + COL_O + static_cast<uint32_t>(t * 128);
    uint8_t* prow = sP + t * P_BYTES + r * 128;
    float m_used = -INFINITY, l = 0.f;
    for (int j = 0; j < n_tiles; ++j) {
      ptx::mbar_wait(&s_full[t * 2 + (j & 1)], (j >> 1) & 1);
      ptx::tc_fence_after();
      uint32_t sreg[BKV / 32][32];
      const uint32_t s_addr = tmem_base + lane_addr + COL_S + static_cast<uint32_t>(t * 128 + (j & 1) * BKV);
#pragma unroll
      for (int c = 0; c < BKV / 32; ++c) ptx::tmem_ld32(s_addr + c * 32, sreg[c]);
      ptx::tmem_ld_wait();
      const int n_valid = S - j * BKV;  // keys >= n_valid are padding (only ever true for the last tile)
      if (n_valid < BKV) {
#pragma unroll
        for (int c = 0; c < BKV / 32; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i >= n_valid) sreg[c][i] = 0xff800000u;  // -inf
      }
      float mxs[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // 4 independent chains
#pragma unroll
      for (int c = 0; c < BKV / 32; ++c)
#pragma unroll
        for (int i = 0; i < 32; ++i) mxs[i & 3] = fmaxf(mxs[i & 3], __uint_as_float(sreg[c][i]));
      const float mx = fmaxf(fmaxf(mxs[0], mxs[1]), fmaxf(mxs[2], mxs[3]));
      const float m_new = fmaxf(m_used, mx * scale_log2);  // scale > 0: max commutes with the scaling
      const bool need = __any_sync(0xffffffffu, m_new > m_used + RESCALE_THRESHOLD);
      float alpha = 1.f;
      if (need) {
        alpha = (m_used == -INFINITY) ? 0.f : exp2f(m_used - m_new);
        m_used = m_new;
      }
      // p = 2^(s*scale - m): one FFMA + one MUFU per element; arguments are <= 8 by construction
      const float neg_m = -m_used;
      float ps[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < BKV / 32; ++c)
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float p = ptx::ex2_approx(fmaf(__uint_as_float(sreg[c][i]), scale_log2, neg_m));
          ps[i & 3] += p;
          sreg[c][i] = __float_as_uint(p);
        }
      l = l * alpha + ((ps[0] + ps[1]) + (ps[2] + ps[3]));
      if (j > 0) {
        // PV of this tile's previous KV tile must be complete: the P buffer is free and O is quiescent
        ptx::mbar_wait(&p_empty[t], (j - 1) & 1);
        ptx::tc_fence_after();
        if (need) {
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t o[32];
            ptx::tmem_ld32(o_addr + c * 32, o);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            ptx::tmem_st32(o_addr + c * 32, o);
          }
          ptx::tmem_st_wait();
        }
      }
      // P (bf16) -> smem, K-major [128 rows][64 keys], 128-byte swizzle: 16-B chunk index ^= row & 7
#pragma unroll
      for (int g = 0; g < BKV / 8; ++g) {
        const int c = g >> 2, i0 = (g & 3) * 8;
        uint4 pk;
        __nv_bfloat162 t0 = __floats2bfloat162_rn(__uint_as_float(sreg[c][i0]), __uint_as_float(sreg[c][i0 + 1]));
        __nv_bfloat162 t1 = __floats2bfloat162_rn(__uint_as_float(sreg[c][i0 + 2]), __uint_as_float(sreg[c][i0 + 3]));
        __nv_bfloat162 t2 = __floats2bfloat162_rn(__uint_as_float(sreg[c][i0 + 4]), __uint_as_float(sreg[c][i0 + 5]));
        __nv_bfloat162 t3 = __floats2bfloat162_rn(__uint_as_float(sreg[c][i0 + 6]), __uint_as_float(sreg[c][i0 + 7]));
        pk.x = *reinterpret_cast<uint32_t*>(&t0);
        pk.y = *reinterpret_cast<uint32_t*>(&t1);
        pk.z = *reinterpret_cast<uint32_t*>(&t2);
        pk.w = *reinterpret_cast<uint32_t*>(&t3);
        *reinterpret_cast<uint4*>(prow + ((g ^ (r & 7)) << 4)) = pk;
      }
      ptx::fence_proxy_async();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&p_full[t]);
    }
    ptx::mbar_wait(o_full, 0);
    ptx::tc_fence_after();
    const int tok = q0 + t * BQ + r;
    const float inv = 1.0f / l;
    bf16* dst = nullptr;
    if (tok < S)
      dst = (tok < Np) ? out_p + (static_cast<long long>(b) * Np + tok) * d + h * HD
                       : out_c + (static_cast<long long>(b) * (S - Np) + (tok - Np)) * d + h * HD;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t o[32];
      ptx::tmem_ld32(o_addr + c * 32, o);
      ptx::tmem_ld_wait();
      if (dst != nullptr) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 pk;
          __nv_bfloat162 t0 = __floats2bfloat162_rn(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv);
          __nv_bfloat162 t1 = __floats2bfloat162_rn(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv);
          __nv_bfloat162 t2 = __floats2bfloat162_rn(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv);
          __nv_bfloat162 t3 = __floats2bfloat162_rn(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv);
          pk.x = *reinterpret_cast<uint32_t*>(&t0);
          pk.y = *reinterpret_cast<uint32_t*>(&t1);
          pk.z = *reinterpret_cast<uint32_t*>(&t2);
          pk.w = *reinterpret_cast<uint32_t*>(&t3);
          *reinterpret_cast<uint4*>(dst + c * 32 + i) = pk;
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc<TMEM_COLS>(tmem_base);
}

}  // namespace

int attention_bf16(const bf16* qkv, int B, int S, int heads, int head_dim, bf16* out_p, int Np, bf16* out_c,
                   cudaStream_t s) {
  LC_REQUIRE(head_dim == HD, "attention: head_dim must be 128");
  static bool attr_set = false;
  if (!attr_set) {
    LC_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  const int d = heads * HD;
  CUtensorMap tmq, tmkv;
  LC_TRY(make_tmap_2d_bf16(&tmq, qkv, static_cast<uint64_t>(3) * d, static_cast<uint64_t>(B) * S,
                           static_cast<uint64_t>(3) * d * 2, 64, 128));
  LC_TRY(make_tmap_2d_bf16(&tmkv, qkv, static_cast<uint64_t>(3) * d, static_cast<uint64_t>(B) * S,
                           static_cast<uint64_t>(3) * d * 2, 64, BKV));
  dim3 grid(ceil_div(S, 2 * BQ), heads, B);
  prof_begin(PROF_ATTN, s);
  attention_tc_kernel<<<grid, NUM_THREADS, SMEM_BYTES, s>>>(tmq, tmkv, S, heads, out_p, Np, out_c);
  prof_end(PROF_ATTN, 4.0 * B * heads * static_cast<double>(S) * S * HD, s);
  LC_LAUNCH_CHECK();
  return 0;
}

}  // namespace lc
