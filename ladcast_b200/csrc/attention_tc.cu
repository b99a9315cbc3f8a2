// tcgen05 flash attention, head_dim 128, non-causal, no mask (F.scaled_dot_product_attention at
// LaDCast_3D_model.py:199-201).  Q/K/V are read by TMA straight out of the token-major [B*S, 3d] projection
// buffer (q | k | v), so no head-major transposes exist anywhere.
//
//   CTA = one (sample, head, 128-query tile), KV tiles of 64 keys; 192 threads, TWO CTAs per SM (96 KB smem and
//   256 TMEM columns each) so that one CTA's softmax overlaps the other's tensor-core work:
//     warp 0 : TMA producer (Q once; K_j / V_j double buffered)
//     warp 1 : TMEM allocator + single-thread tcgen05.mma issuer:  S_j = Q K_j^T  (TMEM, double buffered),
//              O += P_j V_j  (V consumed MN-major, i.e. exactly as it lies in memory)
//     warps 2-5 : softmax; one thread per query row reads its S row from TMEM (no shuffles), online softmax in the
//              log2 domain with lazy rescaling of O (only when the row max grows by > 2^8), writes P_j (bf16) into a
//              128B-swizzled smem tile that is the A operand of the PV product (the tile re-uses K_j's buffer, which is
//              dead once S_j exists); final O / l epilogue.
#include "kernels.h"
#include "ptx.cuh"
#include "tmap.h"

namespace lc {
namespace {

constexpr int HD = 128;
constexpr int BQ = 128;
constexpr int BKV = 64;
constexpr int Q_BYTES = 128 * 128 * 2;     // 32 KB: two [128 rows][64 dims] swizzled boxes
constexpr int QSUB_BYTES = 128 * 64 * 2;   // 16 KB
constexpr int KV_BYTES = BKV * 128 * 2;    // 16 KB: two [64 keys][64 dims] swizzled boxes
constexpr int KVSUB_BYTES = BKV * 64 * 2;  // 8 KB
constexpr int SMEM_BYTES = Q_BYTES + 2 * KV_BYTES /*K (+P alias)*/ + 2 * KV_BYTES /*V*/ + 1024 + 256;
constexpr int NUM_THREADS = 192;
constexpr uint32_t TMEM_COLS = 256;
constexpr uint32_t COL_S0 = 0, COL_O = 128;
constexpr float RESCALE_THRESHOLD = 8.0f;

__global__ void __launch_bounds__(NUM_THREADS, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmkv, int S, int heads, bf16* __restrict__ out_p, int Np,
                    bf16* __restrict__ out_c) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Q_BYTES;
  uint8_t* sV = sK + 2 * KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * KV_BYTES);
  uint64_t* q_full = bars;           // 1
  uint64_t* k_full = bars + 1;       // 2
  uint64_t* v_full = bars + 3;       // 2
  uint64_t* kv_empty = bars + 5;     // 2
  uint64_t* s_full = bars + 7;       // 2
  uint64_t* p_full = bars + 9;       // 2 (4 arrivals each: one per softmax warp); P_j sits in K stage j&1
  uint64_t* p_empty = bars + 11;     // 1 (PV_j complete; only waited on when O has to be rescaled)
  uint64_t* o_full = bars + 12;      // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d = heads * HD;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * BQ;
  const int n_tiles = (S + BKV - 1) / BKV;
  const int row_base = b * S;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmq);
    ptx::prefetch_tmap(&tmkv);
    ptx::mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&k_full[i], 1);
      ptx::mbar_init(&v_full[i], 1);
      ptx::mbar_init(&kv_empty[i], 1);
      ptx::mbar_init(&s_full[i], 1);
    }
    ptx::mbar_init(&p_full[0], 4);
    ptx::mbar_init(&p_full[1], 4);
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
    ptx::mbar_expect_tx(q_full, Q_BYTES);
    ptx::tma_load_2d(sQ, &tmq, q_full, h * HD, row_base + q0);
    ptx::tma_load_2d(sQ + QSUB_BYTES, &tmq, q_full, h * HD + 64, row_base + q0);
    for (int j = 0; j < n_tiles; ++j) {
      const int st = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      ptx::mbar_wait(&kv_empty[st], ph ^ 1);
      const int kr = row_base + j * BKV;
      ptx::mbar_expect_tx(&k_full[st], KV_BYTES);
      ptx::tma_load_2d(sK + st * KV_BYTES, &tmkv, &k_full[st], d + h * HD, kr);
      ptx::tma_load_2d(sK + st * KV_BYTES + KVSUB_BYTES, &tmkv, &k_full[st], d + h * HD + 64, kr);
      ptx::mbar_expect_tx(&v_full[st], KV_BYTES);
      ptx::tma_load_2d(sV + st * KV_BYTES, &tmkv, &v_full[st], 2 * d + h * HD, kr);
      ptx::tma_load_2d(sV + st * KV_BYTES + KVSUB_BYTES, &tmkv, &v_full[st], 2 * d + h * HD + 64, kr);
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_qk = ptx::make_idesc_bf16(128, BKV, 0, 0);  // A = Q (K-major), B = K (K-major)
    constexpr uint32_t idesc_pv = ptx::make_idesc_bf16(128, 128, 0, 1);  // A = P (K-major), B = V (MN-major)
    const uint32_t q_addr = ptx::smem_u32(sQ);
    auto issue_qk = [&](int j) {
      const int st = j & 1;
      ptx::mbar_wait(&k_full[st], (j >> 1) & 1);
      ptx::tc_fence_after();
      const uint32_t k_addr = ptx::smem_u32(sK + st * KV_BYTES);
      const uint32_t d_tmem = tmem_base + COL_S0 + static_cast<uint32_t>(st * BKV);
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {  // 128 head dims = 2 sub-tiles x 4 K-steps of 16
        const uint32_t qoff = (kk >> 2) * QSUB_BYTES + (kk & 3) * 32;
        const uint32_t koff = (kk >> 2) * KVSUB_BYTES + (kk & 3) * 32;
        ptx::umma_f16(d_tmem, ptx::make_smem_desc(q_addr + qoff, 16, 1024), ptx::make_smem_desc(k_addr + koff, 16, 1024),
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
      ptx::mbar_wait(&p_full[st], (j >> 1) & 1);
      ptx::tc_fence_after();
      const uint32_t v_addr = ptx::smem_u32(sV + st * KV_BYTES);
      const uint32_t p_addr = ptx::smem_u32(sK + st * KV_BYTES);  // P_j lives in K_j's (dead) buffer
#pragma unroll
      for (int kk = 0; kk < BKV / 16; ++kk) {  // 64 keys = 4 K-steps of 16
        const uint64_t da = ptx::make_smem_desc(p_addr + kk * 32, 16, 1024);
        const uint64_t db = ptx::make_smem_desc(v_addr + kk * 2048, KVSUB_BYTES, 1024);
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
      uint32_t sreg[BKV / 32][32];
      const uint32_t s_addr = tmem_base + lane_addr + COL_S0 + static_cast<uint32_t>(st * BKV);
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
      float ps[4] = {0.f, 0.f, 0.f, 0.f};  // independent partial sums: no 64-long dependent FADD chain
#pragma unroll
      for (int c = 0; c < BKV / 32; ++c)
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float p = ptx::ex2_approx(fmaf(__uint_as_float(sreg[c][i]), scale_log2, neg_m));
          ps[i & 3] += p;
          sreg[c][i] = __float_as_uint(p);
        }
      l = l * alpha + ((ps[0] + ps[1]) + (ps[2] + ps[3]));
      // P_j goes to K stage j&1, which nothing else touches until PV_j: no wait needed to write it.  Only a
      // rescale of O has to wait for PV_{j-1} (rare after the first tiles thanks to the 2^8 lazy threshold).
      if (j > 0 && need) {
        ptx::mbar_wait(p_empty, (j - 1) & 1);
        ptx::tc_fence_after();
        {
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
      // P (bf16) -> smem (K_j's buffer), K-major [128 rows][64 keys], 128-byte swizzle: 16-B chunk index ^= row & 7
      uint8_t* prow = sK + st * KV_BYTES + r * 128;
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
      if (lane == 0) ptx::mbar_arrive(&p_full[st]);
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
  CUtensorMap tmq, tmkv;
  LC_TRY(make_tmap_2d_bf16(&tmq, qkv, static_cast<uint64_t>(3) * d, static_cast<uint64_t>(B) * S,
                           static_cast<uint64_t>(3) * d * 2, 64, 128));
  LC_TRY(make_tmap_2d_bf16(&tmkv, qkv, static_cast<uint64_t>(3) * d, static_cast<uint64_t>(B) * S,
                           static_cast<uint64_t>(3) * d * 2, 64, BKV));
  dim3 grid(ceil_div(S, BQ), heads, B);
  prof_begin(PROF_ATTN, s);
  attention_tc_kernel<<<grid, NUM_THREADS, SMEM_BYTES, s>>>(tmq, tmkv, S, heads, out_p, Np, out_c);
  prof_end(PROF_ATTN, 4.0 * B * heads * static_cast<double>(S) * S * HD, s);
  LC_LAUNCH_CHECK();
  return 0;
}

}  // namespace lc
