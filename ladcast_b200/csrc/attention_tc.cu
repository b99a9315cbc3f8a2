// tcgen05 flash attention, head_dim 128, non-causal, no mask (F.scaled_dot_product_attention at
// LaDCast_3D_model.py:199-201).  Q/K/V are read by TMA straight out of the token-major [B*S, 3d] projection
// buffer (q | k | v), so no head-major transposes exist anywhere.
//
// Shape of the kernel follows from the measured tcgen05.mma issue rate (tools/ubench/mma_rate.cu): an M=128 MMA costs
// ~75-80 cycles whether N is 64 or 128, so S tiles are 128 keys wide, and reading the A operand from tensor memory
// is the fastest form, so P never goes through shared memory:
//
//   CTA = one (sample, head, 256-query block) = two 128-query tiles A and B that share every K/V tile; 576 threads:
//     warp 0    : TMA producer — Q (both tiles) once, then K_0 V_0 K_1 V_1 ... (128 keys x 128 dims = 32 KB each)
//                 through a 5-slot ring, ~2.5 KV steps ahead of the tensor pipe
//     warp 1    : TMEM allocator + single-thread tcgen05.mma issuer.  Per tile and KV step: O_t += P_t V_j with P_t
//                 read from TMEM, then S_t = Q_t K_{j+1}^T into the same TMEM columns (S and the bf16 P alias; the
//                 tensor pipe executes in issue order, so the overwrite is safe).  The two tiles ping-pong: while
//                 one tile's softmax runs, the tensor pipe works for the other.
//     warps 2-9 : softmax of tile A, warps 10-17: softmax of tile B.  Two threads per query row (64 keys each, kept
//                 in registers) read S from TMEM (no shuffles), complete the row max through a 2 KB shared-memory
//                 exchange + a 64-thread named barrier, online softmax in the log2 domain (one FFMA + one MUFU.EX2
//                 per element), lazy rescale of O (only when the row max grows by more than 2^8), P (bf16) stored
//                 back to TMEM over S.  Final O / l epilogue per tile.  576 threads leave 96 registers per thread.
//   TMEM columns: S_A/P_A 0..127, S_B/P_B 128..255, O_A 256..383, O_B 384..511.
#include "kernels.h"
#include "ptx.cuh"
#include "tmap.h"

namespace lc {
namespace {

constexpr int HD = 128;
constexpr int BQ = 128;                     // rows per query tile; two tiles per CTA
constexpr int BKV = 128;
constexpr int RING = 5;
constexpr int TILE_BYTES = 128 * 128 * 2;   // 32 KB: two [128 rows][64 dims] swizzled boxes
constexpr int SUB_BYTES = 128 * 64 * 2;     // 16 KB
constexpr int XCH_BYTES = 2 * 2 * 128 * 4;  // row-max exchange between the two threads of a row: [tile][half][row]
// no alignment slack: the kernel has no static shared memory, so the dynamic window starts 1024-aligned (checked)
constexpr int SMEM_BYTES = 2 * TILE_BYTES + RING * TILE_BYTES + 256 + XCH_BYTES;
constexpr int NUM_THREADS = 576;  // TMA warp, MMA warp, 2 tiles x 8 softmax warps
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t COL_S = 0;    // S[tile] (fp32, 128 cols) / P[tile] (bf16 pairs, first 64 cols) at COL_S + tile*128
constexpr uint32_t COL_O = 256;  // O[tile] at COL_O + tile*128
constexpr float RESCALE_THRESHOLD = 8.0f;

// Optional in-kernel timeline of CTA (0,0,0) for tuning (tools/attn_trace.py): clock64 stamps, [role][step][4].
__device__ long long* g_trace = nullptr;
#define LC_TRACE(role, j, k)                                           \
  do {                                                                 \
    if (trace != nullptr) trace[((role) * 64 + (j)) * 4 + (k)] = clock64(); \
  } while (0)

__global__ void __maxnreg__(96)
attention_tc_kernel(const __grid_constant__ CUtensorMap tm, int S, int heads, bf16* __restrict__ out_p, int Np,
                    bf16* __restrict__ out_c) {
  extern __shared__ uint8_t smem_raw[];
  if ((ptx::smem_u32(smem_raw) & 1023u) != 0) __trap();  // swizzled TMA tiles need 1024-byte alignment
  uint8_t* smem = smem_raw;
  uint8_t* sQ = smem;                        // [2][TILE_BYTES]
  uint8_t* sR = sQ + 2 * TILE_BYTES;         // [RING][TILE_BYTES]: K_0 V_0 K_1 V_1 ...
  uint64_t* bars = reinterpret_cast<uint64_t*>(sR + RING * TILE_BYTES);
  uint64_t* q_full = bars;                   // 1
  uint64_t* r_full = bars + 1;               // RING
  uint64_t* r_empty = r_full + RING;         // RING
  uint64_t* s_full = r_empty + RING;         // [tile]: S_t(j) complete (and with it every earlier MMA)
  uint64_t* p_full = s_full + 2;             // [tile] (8 arrivals: one per softmax warp)
  uint64_t* o_full = p_full + 2;             // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);
  float* xch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d = heads * HD;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * (2 * BQ);
  const int n_tiles = (S + BKV - 1) / BKV;
  const int row_base = b * S;
  long long* trace = (blockIdx.x | blockIdx.y | blockIdx.z) == 0 && lane == 0 ? g_trace : nullptr;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm);
    ptx::mbar_init(q_full, 1);
    for (int i = 0; i < RING; ++i) {
      ptx::mbar_init(&r_full[i], 1);
      ptx::mbar_init(&r_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&s_full[i], 1);
      ptx::mbar_init(&p_full[i], 8);
    }
    ptx::mbar_init(o_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<TMEM_COLS>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_grid_sync();  // q|k|v are produced by the preceding kernel: nothing global is touched above this line

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    ptx::mbar_expect_tx(q_full, 2 * TILE_BYTES);
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      ptx::tma_load_2d(sQ + t * TILE_BYTES, &tm, q_full, h * HD, row_base + q0 + t * BQ);
      ptx::tma_load_2d(sQ + t * TILE_BYTES + SUB_BYTES, &tm, q_full, h * HD + 64, row_base + q0 + t * BQ);
    }
    int st = 0;
    uint32_t ph = 0;
    for (int j = 0; j < n_tiles; ++j) {
      const int kr = row_base + j * BKV;
#pragma unroll
      for (int kv = 0; kv < 2; ++kv) {  // 0: K_j, 1: V_j
        ptx::mbar_wait(&r_empty[st], ph ^ 1);
        const int col = (kv + 1) * d + h * HD;
        uint8_t* dst = sR + st * TILE_BYTES;
        ptx::mbar_expect_tx(&r_full[st], TILE_BYTES);
        ptx::tma_load_2d(dst, &tm, &r_full[st], col, kr);
        ptx::tma_load_2d(dst + SUB_BYTES, &tm, &r_full[st], col + 64, kr);
        if (++st == RING) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_qk = ptx::make_idesc_bf16(128, BKV, 0, 0);  // A = Q (K-major), B = K (K-major)
    constexpr uint32_t idesc_pv = ptx::make_idesc_bf16(128, HD, 0, 1);   // A = P (TMEM), B = V (MN-major)
    const uint32_t q_addr = ptx::smem_u32(sQ), r_addr = ptx::smem_u32(sR);
    auto issue_qk = [&](int t, int slot) {
      const uint32_t k_addr = r_addr + slot * TILE_BYTES;
      const uint32_t d_tmem = tmem_base + COL_S + static_cast<uint32_t>(t * 128);
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {  // 128 head dims = 2 sub-tiles x 4 K-steps of 16
        const uint32_t off = (kk >> 2) * SUB_BYTES + (kk & 3) * 32;
        ptx::umma_f16(d_tmem, ptx::make_smem_desc(q_addr + t * TILE_BYTES + off, 16, 1024),
                      ptx::make_smem_desc(k_addr + off, 16, 1024), idesc_qk, kk != 0 ? 1u : 0u);
      }
      ptx::umma_commit(&s_full[t]);
    };
    auto issue_pv = [&](int t, int slot, bool first) {
      const uint32_t v_addr = r_addr + slot * TILE_BYTES;
      const uint32_t d_tmem = tmem_base + COL_O + static_cast<uint32_t>(t * 128);
      const uint32_t a_tmem = tmem_base + COL_S + static_cast<uint32_t>(t * 128);
#pragma unroll
      for (int kk = 0; kk < BKV / 16; ++kk)  // 128 keys = 8 K-steps of 16 (16 key rows x 128 B = 2 KB per sub-tile)
        ptx::umma_f16_ts(d_tmem, a_tmem + kk * 8, ptx::make_smem_desc(v_addr + kk * 2048, SUB_BYTES, 1024), idesc_pv,
                         (first && kk == 0) ? 0u : 1u);
    };
    int st = 0;        // ring slot of the next tile to consume (K_0 first)
    uint32_t ph = 0;
    auto advance = [&]() { if (++st == RING) { st = 0; ph ^= 1; } };
    ptx::mbar_wait(q_full, 0);
    ptx::mbar_wait(&r_full[st], ph);
    ptx::tc_fence_after();
    issue_qk(0, st);
    issue_qk(1, st);
    ptx::umma_commit(&r_empty[st]);
    advance();
    for (int j = 0; j < n_tiles; ++j) {
      const int v_slot = st;
      ptx::mbar_wait(&r_full[v_slot], ph);
      advance();
      const int k_slot = st;  // K_{j+1}
      const uint32_t k_ph = ph;
      const bool more = j + 1 < n_tiles;
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        ptx::mbar_wait(&p_full[t], j & 1);
        ptx::tc_fence_after();
        LC_TRACE(0, j, 2 * t);
        issue_pv(t, v_slot, j == 0);
        if (t == 1) ptx::umma_commit(&r_empty[v_slot]);
        if (more) {
          if (t == 0) {
            ptx::mbar_wait(&r_full[k_slot], k_ph);
            ptx::tc_fence_after();
          }
          issue_qk(t, k_slot);
          if (t == 1) ptx::umma_commit(&r_empty[k_slot]);
        }
        LC_TRACE(0, j, 2 * t + 1);
      }
      if (more) advance();
    }
    ptx::umma_commit(o_full);
  } else if (warp >= 2) {
    // ===================== softmax + epilogue =====================
    // Two threads per query row (warps with the same TMEM lane quarter): each owns 64 of the row's 128 keys, which
    // shortens the per-step softmax latency (it sits in the serial softmax -> PV -> QK chain of a tile): a lone warp
    // per sub-partition issues one MUFU.EX2 per ~14 cycles, two warps together one per ~11.5 (tools/ubench/exp_rate.cu).
    const int idx = warp - 2;
    const int t = idx >> 3;          // 0: tile A (warps 2-9), 1: tile B (warps 10-17)
    const int hh = (idx >> 2) & 1;   // key half of the row this thread exponentiates
    const int quarter = warp & 3;    // TMEM lane quarter this warp may access (= warp id % 4)
    if (idx & 7) trace = nullptr;    // one traced warp per tile
    const int r = quarter * 32 + lane;
    const uint32_t pair_bar = 1 + t * 4 + quarter;  // named barrier of the two warps sharing these 32 rows
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const float scale_log2 = 0.08838834764831845f * 1.4426950408889634f;
    const uint32_t s_addr = tmem_base + lane_addr + COL_S + static_cast<uint32_t>(t * 128);
    const uint32_t o_addr = tmem_base + lane_addr + COL_O + static_cast<uint32_t>(t * 128 + hh * 64);
    float m_used = -INFINITY, l = 0.f;
    for (int j = 0; j < n_tiles; ++j) {
      // S_t(j) complete; the commit also covers PV_t(j-1), so O_t is quiescent until this tile hands over P_t(j)
      LC_TRACE(1 + t, j, 0);
      ptx::mbar_wait(&s_full[t], j & 1);
      ptx::tc_fence_after();
      LC_TRACE(1 + t, j, 1);
      const int n_valid = S - j * BKV;  // keys >= n_valid are padding (only ever true for the last tile)
      // this thread's 64 keys of the row stay in registers; the row max is completed with the partner thread (the
      // other warp on the same TMEM lanes) through shared memory.  The pair barrier doubles as "both halves of S
      // have been read", after which P may overwrite the S columns.
      uint32_t sreg[2][32];
      ptx::tmem_ld32(s_addr + hh * 64, sreg[0]);
      ptx::tmem_ld32(s_addr + hh * 64 + 32, sreg[1]);
      ptx::tmem_ld_wait();
      if (n_valid < BKV) {
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (hh * 64 + c * 32 + i >= n_valid) sreg[c][i] = 0xff800000u;  // -inf
      }
      float mxs[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // 4 independent chains
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int i = 0; i < 32; ++i) mxs[i & 3] = fmaxf(mxs[i & 3], __uint_as_float(sreg[c][i]));
      float mx = fmaxf(fmaxf(mxs[0], mxs[1]), fmaxf(mxs[2], mxs[3]));
      xch[(t * 2 + hh) * 128 + r] = mx;
      asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
      mx = fmaxf(mx, xch[(t * 2 + (1 - hh)) * 128 + r]);
      LC_TRACE(1 + t, j, 2);
      const float m_new = fmaxf(m_used, mx * scale_log2);  // scale > 0: max commutes with the scaling
      const bool need = __any_sync(0xffffffffu, m_new > m_used + RESCALE_THRESHOLD);
      float alpha = 1.f;
      if (need) {
        alpha = (m_used == -INFINITY) ? 0.f : exp2f(m_used - m_new);
        m_used = m_new;
      }
      // p = 2^(s*scale - m): one FFMA + one MUFU per element; arguments are <= 8 by construction.
      // P (bf16 pairs) goes back to TMEM over S: columns [32*hh, 32*hh+32) hold keys [64*hh, 64*hh+64).
      const float neg_m = -m_used;
      float ps[4] = {0.f, 0.f, 0.f, 0.f};
      uint32_t pk[32];
#pragma unroll
      for (int i = 0; i < 64; i += 2) {
        const float p0 = ptx::ex2_approx(fmaf(__uint_as_float(sreg[i >> 5][i & 31]), scale_log2, neg_m));
        const float p1 = ptx::ex2_approx(fmaf(__uint_as_float(sreg[i >> 5][(i & 31) + 1]), scale_log2, neg_m));
        ps[(i >> 1) & 3] += p0 + p1;
        __nv_bfloat162 pb = __floats2bfloat162_rn(p0, p1);
        pk[i >> 1] = *reinterpret_cast<uint32_t*>(&pb);
      }
      if (need && j > 0) {  // rare: rescale this thread's 64 columns of O (quiescent: PV_t(j-1) has completed)
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t o[32];
          ptx::tmem_ld32(o_addr + c * 32, o);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          ptx::tmem_st32(o_addr + c * 32, o);
        }
      }
      ptx::tmem_st32(s_addr + hh * 32, pk);
      l = l * alpha + ((ps[0] + ps[1]) + (ps[2] + ps[3]));
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&p_full[t]);
      LC_TRACE(1 + t, j, 3);
    }
    ptx::mbar_wait(o_full, 0);
    ptx::tc_fence_after();
    // row sum = the two halves' partial sums (same alpha sequence in both); Q's shared memory is dead by now
    float* lx = reinterpret_cast<float*>(sQ);
    lx[(t * 2 + hh) * 128 + r] = l;
    asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
    const float inv = 1.0f / (l + lx[(t * 2 + (1 - hh)) * 128 + r]);
    const int tok = q0 + t * BQ + r;
    bf16* dst = nullptr;
    if (tok < S)
      dst = ((tok < Np) ? out_p + (static_cast<long long>(b) * Np + tok) * d
                        : out_c + (static_cast<long long>(b) * (S - Np) + (tok - Np)) * d) + h * HD + hh * 64;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
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

int attention_set_trace(long long* buf) {
  LC_CHECK_CUDA(cudaMemcpyToSymbol(g_trace, &buf, sizeof(buf)));
  return 0;
}

int attention_bf16(const bf16* qkv, int B, int S, int heads, int head_dim, bf16* out_p, int Np, bf16* out_c,
                   cudaStream_t s) {
  LC_REQUIRE(head_dim == HD, "attention: head_dim must be 128");
  static PerDevice<bool> attr_set;
  if (!attr_set.here()) {
    LC_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set.here() = true;
  }
  const int d = heads * HD;
  CUtensorMap tm;
  LC_TRY(make_tmap_2d_bf16(&tm, qkv, static_cast<uint64_t>(3) * d, static_cast<uint64_t>(B) * S,
                           static_cast<uint64_t>(3) * d * 2, 64, 128));
  dim3 grid(ceil_div(S, 2 * BQ), heads, B);
  prof_begin(PROF_ATTN, s);
  LC_CHECK_CUDA(launch_kernel(attention_tc_kernel, grid, NUM_THREADS, SMEM_BYTES, s, tm, S, heads, out_p, Np, out_c));
  prof_end(PROF_ATTN, 4.0 * B * heads * static_cast<double>(S) * S * HD, s, 8.0 * B * S * d);  // q, k, v in + o out (bf16)
  LC_LAUNCH_CHECK();
  return 0;
}

}  // namespace lc
