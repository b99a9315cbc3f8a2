// tcgen05 flash attention, head_dim 128, non-causal, no mask (F.scaled_dot_product_attention at
// LaDCast_3D_model.py:199-201).  Q/K/V are read by TMA straight out of the token-major [B*S, 3d] projection
// buffer (q | k | v), so no head-major transposes exist anywhere.
//
//   CTA = one (sample, head, 128-query tile); 192 threads:
//     warp 0 : TMA producer (Q once; K_j / V_j double buffered)
//     warp 1 : TMEM allocator + single-thread tcgen05.mma issuer:  S_j = Q K_j^T  (TMEM, double buffered),
//              O += P_j V_j  (V consumed MN-major, i.e. exactly as it lies in memory)
//     warps 2-5 : softmax; one thread per query row reads its S row from TMEM (no shuffles), online softmax in the
//              log2 domain with lazy rescaling of O (only when the row max grows by > 2^8), writes P_j (bf16) into a
//              128B-swizzled smem tile that is the A operand of the PV product; final O / l epilogue.
#include "kernels.h"
#include "ptx.cuh"
#include "tmap.h"

namespace lc {
namespace {

constexpr int HD = 128;
constexpr int BQ = 128;
constexpr int BKV = 128;
constexpr int TILE_BYTES = 128 * 128 * 2;  // 32 KB: two [128 rows][64 cols] swizzled boxes
constexpr int SUB_BYTES = 128 * 64 * 2;    // 16 KB
constexpr int SMEM_BYTES = TILE_BYTES /*Q*/ + 2 * TILE_BYTES /*K*/ + 2 * TILE_BYTES /*V*/ + TILE_BYTES /*P*/ + 1024 + 256;
constexpr int NUM_THREADS = 192;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t COL_S0 = 0, COL_O = 256;
constexpr float RESCALE_THRESHOLD = 8.0f;

__global__ void __launch_bounds__(NUM_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tm, int S, int heads, bf16* __restrict__ out_p, int Np,
                    bf16* __restrict__ out_c) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + TILE_BYTES;
  uint8_t* sV = sK + 2 * TILE_BYTES;
  uint8_t* sP = sV + 2 * TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + TILE_BYTES);
  uint64_t* q_full = bars;           // 1
  uint64_t* k_full = bars + 1;       // 2
  uint64_t* v_full = bars + 3;       // 2
  uint64_t* kv_empty = bars + 5;     // 2
  uint64_t* s_full = bars + 7;       // 2
  uint64_t* p_full = bars + 9;       // 1 (4 arrivals: one per softmax warp)
  uint64_t* p_empty = bars + 10;     // 1
  uint64_t* o_full = bars + 11;      // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d = heads * HD;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * BQ;
  const int n_tiles = (S + BKV - 1) / BKV;
  const int row_base = b * S;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm);
    ptx::mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&k_full[i], 1);
      ptx::mbar_init(&v_full[i], 1);
      ptx::mbar_init(&kv_empty[i], 1);
      ptx::mbar_init(&s_full[i], 1);
    }
    ptx::mbar_init(p_full, 4);
    ptx::mbar_init(p_empty, 1);
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
    ptx::mbar_expect_tx(q_full, TILE_BYTES);
    ptx::tma_load_2d(sQ, &tm, q_full, h * HD, row_base + q0);
    ptx::tma_load_2d(sQ + SUB_BYTES, &tm, q_full, h * HD + 64, row_base + q0);
    for (int j = 0; j < n_tiles; ++j) {
      const int st = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      ptx::mbar_wait(&kv_empty[st], ph ^ 1);
      const int kr = row_base + j * BKV;
      ptx::mbar_expect_tx(&k_full[st], TILE_BYTES);
      ptx::tma_load_2d(sK + st * TILE_BYTES, &tm, &k_full[st], d + h * HD, kr);
      ptx::tma_load_2d(sK + st * TILE_BYTES + SUB_BYTES, &tm, &k_full[st], d + h * HD + 64, kr);
      ptx::mbar_expect_tx(&v_full[st], TILE_BYTES);
      ptx::tma_load_2d(sV + st * TILE_BYTES, &tm, &v_full[st], 2 * d + h * HD, kr);
      ptx::tma_load_2d(sV + st * TILE_BYTES + SUB_BYTES, &tm, &v_full[st], 2 * d + h * HD + 64, kr);
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_qk = ptx::make_idesc_bf16(128, 128, 0, 0);  // A = Q (K-major), B = K (K-major)
    constexpr uint32_t idesc_pv = ptx::make_idesc_bf16(128, 128, 0, 1);  // A = P (K-major), B = V (MN-major)
    const uint32_t q_addr = ptx::smem_u32(sQ), p_addr = ptx::smem_u32(sP);
    auto issue_qk = [&](int j) {
      const int st = j & 1;
      ptx::mbar_wait(&k_full[st], (j >> 1) & 1);
      ptx::tc_fence_after();
      const uint32_t k_addr = ptx::smem_u32(sK + st * TILE_BYTES);
      const uint32_t d_tmem = tmem_base + COL_S0 + static_cast<uint32_t>(st * 128);
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint32_t off = (kk >> 2) * SUB_BYTES + (kk & 3) * 32;
        ptx::umma_f16(d_tmem, ptx::make_smem_desc(q_addr + off, 16, 1024), ptx::make_smem_desc(k_addr + off, 16, 1024),
                      idesc_qk, kk != 0 ? 1u : 0u);
      }
      ptx::umma_commit(&s_full[st]);
    };
    ptx::mbar_wait(q_full, 0);
    issue_qk(0);
    for (int j = 0; j < n_tiles; ++j) {
      const int st = j & 1;
      if (j + 1 < n_tiles) issue_qk(j + 1);
      ptx::mbar_wait(&v_full[st], (j >> 1) & 1);
      ptx::mbar_wait(p_full, j & 1);
      ptx::tc_fence_after();
      const uint32_t v_addr = ptx::smem_u32(sV + st * TILE_BYTES);
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint64_t da = ptx::make_smem_desc(p_addr + (kk >> 2) * SUB_BYTES + (kk & 3) * 32, 16, 1024);
        const uint64_t db = ptx::make_smem_desc(v_addr + kk * 2048, SUB_BYTES, 1024);
        ptx::umma_f16(tmem_base + COL_O, da, db, idesc_pv, (j | kk) != 0 ? 1u : 0u);
      }
      ptx::umma_commit(&kv_empty[st]);
      ptx::umma_commit(p_empty);
    }
    ptx::umma_commit(o_full);
  } else if (warp >= 2) {
    // ===================== softmax + epilogue (thread = query row) =====================
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const float scale_log2 = 0.08838834764831845f * 1.4426950408889634f;
    float m_used = -INFINITY, l = 0.f;
    for (int j = 0; j < n_tiles; ++j) {
      const int st = j & 1;
      ptx::mbar_wait(&s_full[st], (j >> 1) & 1);
      ptx::tc_fence_after();
      uint32_t sreg[4][32];
      const uint32_t s_addr = tmem_base + lane_addr + COL_S0 + static_cast<uint32_t>(st * 128);
#pragma unroll
      for (int c = 0; c < 4; ++c) ptx::tmem_ld32(s_addr + c * 32, sreg[c]);
      ptx::tmem_ld_wait();
      const int n_valid = S - j * BKV;  // keys >= n_valid are padding
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float v = __uint_as_float(sreg[c][i]) * scale_log2;
          if (c * 32 + i >= n_valid) v = -INFINITY;
          sreg[c][i] = __float_as_uint(v);
          mx = fmaxf(mx, v);
        }
      const float m_new = fmaxf(m_used, mx);
      const bool need = __any_sync(0xffffffffu, m_new > m_used + RESCALE_THRESHOLD);
      float alpha = 1.f;
      if (need) {
        alpha = (m_used == -INFINITY) ? 0.f : exp2f(m_used - m_new);
        m_used = m_new;
      }
      float psum = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float p = exp2f(__uint_as_float(sreg[c][i]) - m_used);
          psum += p;
          sreg[c][i] = __float_as_uint(p);
        }
      l = l * alpha + psum;
      if (j > 0) {
        ptx::mbar_wait(p_empty, (j - 1) & 1);  // PV_{j-1} finished: O is quiescent and the P tile is free
        ptx::tc_fence_after();
        if (need) {
          const uint32_t o_addr = tmem_base + lane_addr + COL_O;
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
      // P (bf16) -> smem, K-major [128 rows][64 keys] x 2, 128-byte swizzle: 16-B chunk index ^= row & 7
      uint8_t* prow = sP + r * 128;
#pragma unroll
      for (int g = 0; g < 16; ++g) {
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
        const int sub = g >> 3, chunk = g & 7;
        *reinterpret_cast<uint4*>(prow + sub * SUB_BYTES + ((chunk ^ (r & 7)) << 4)) = pk;
      }
      ptx::fence_proxy_async();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(p_full);
    }
    ptx::mbar_wait(o_full, 0);
    ptx::tc_fence_after();
    const int tok = q0 + r;
    const float inv = 1.0f / l;
    bf16* dst = nullptr;
    if (tok < S)
      dst = (tok < Np) ? out_p + (static_cast<long long>(b) * Np + tok) * d + h * HD
                       : out_c + (static_cast<long long>(b) * (S - Np) + (tok - Np)) * d + h * HD;
    const uint32_t o_addr = tmem_base + lane_addr + COL_O;
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
  CUtensorMap tm;
  LC_TRY(make_tmap_2d_bf16(&tm, qkv, static_cast<uint64_t>(3) * d, static_cast<uint64_t>(B) * S,
                           static_cast<uint64_t>(3) * d * 2, 64, 128));
  dim3 grid(ceil_div(S, BQ), heads, B);
  prof_begin(PROF_ATTN, s);
  attention_tc_kernel<<<grid, NUM_THREADS, SMEM_BYTES, s>>>(tm, S, heads, out_p, Np, out_c);
  prof_end(PROF_ATTN, 4.0 * B * heads * static_cast<double>(S) * S * HD, s);
  LC_LAUNCH_CHECK();
  return 0;
}

}  // namespace lc
