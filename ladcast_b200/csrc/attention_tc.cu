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
#include <cstdlib>
#include <type_traits>

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
  if (warp == 0) LC_TRACE(3, 0, 0);  // role 3: CTA life cycle — entry, set-up done, O complete, last store issued

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
  if (warp == 0) LC_TRACE(3, 0, 1);

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
    LC_TRACE(3, 0, 2);
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
    LC_TRACE(3, 0, 3);
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc<TMEM_COLS>(tmem_base);
}


// ------------------------------------------------------------------------------------------------ persistent variant
// Same tile algorithm, but ONE CTA per SM walks over (sample, head, 256-query block) work items, so that per-item
// set-up and drain are paid once per CTA or overlap the next item instead of costing ~21 % of every CTA's life
// (in-kernel timeline of the one-item kernel: 62 k cycles of steady-state steps in a 79 k-cycle CTA):
//   * barriers, TMEM and the K/V ring persist; the producer streams the next item's K/V right behind the current one
//     and re-loads Q as soon as the item's last S = Q K^T has completed (q_empty);
//   * the next item's first S tiles are issued straight after the current item's last P V; its first P V waits for
//     the epilogue to drain O (o_empty[tile]);
//   * the epilogue no longer issues 16-byte stores scattered over 128 rows per warp instruction (LSU-bound, ~7.6 k
//     cycles): O / l is packed to bf16 into a swizzled 32 KB staging tile and written by TMA tensor stores (3-D maps
//     {d, tokens of the stream, samples}: rows past the stream's token count are clipped by the hardware, which also
//     splits a tile that straddles the pred | cond boundary).  The staging tiles are slots of the K/V ring: the
//     producer passes the two slots that follow V_{n-1} of every item to the softmax warps instead of filling them
//     (ring order K_0 V_0 ... K_{n-1} V_{n-1} [O_A staging] [O_B staging]) — all 5 slots carry K/V in the steady
//     state and no extra shared memory is needed;
//   * the thread that issues a tile's TMA stores does not wait for them: it returns the staging slot to the producer
//     (cp.async.bulk.wait_group.read + arrive) only after the wait for the next item's first S tile, when the ~1.9 k
//     cycles the stores need to drain the tile have long passed.  (A dedicated 19th store warp was tried instead: the
//     item boundary got 1-2.7 k cycles shorter but every softmax step 180 cycles longer — no net gain.)
// Shared memory: Q 2 x 32 KB, ring 5 x 32 KB (as the one-item kernel).
constexpr int RING_P = 5;
constexpr int ATTN_POLY_DEFAULT = 0;
constexpr int ATTN_SPLIT_DEFAULT = 1;
constexpr int SMEM_BYTES_P = 2 * TILE_BYTES + RING_P * TILE_BYTES + 256 + XCH_BYTES;
constexpr int NUM_THREADS_P = NUM_THREADS;

// POLY > 0: every POLY-th element of a row's exponentials is computed by ptx::ex2_poly on the FMA / ALU pipes instead
// of the MUFU (16 ex2 / clk / SM against 128 x 128 exponentials per tile step).  Measured on B200 (B=20, S=2250, H=12,
// profiles/r02_attn_ab.log): every period tried (2, 3, 4, 8) is SLOWER than the all-MUFU form (1063 -> 976-1048 TFLOP/s;
// with the split hand-over 1114 -> 999-1111): the softmax step is issue- / latency-bound, not MUFU-throughput-bound, and
// the 8 extra FMA / ALU instructions per element cost more than the MUFU slot they free.  Kept as an experiment
// (LADCAST_B200_ATTN_POLY=3), off by default.
// SPLIT: a thread owns keys [32 hh, 32 hh + 32) and [64 + 32 hh, 64 + 32 hh + 32) of its row instead of one 64-key
// half, and P is handed to the tensor pipe in two 64-key chunks (p_half, then p_full): the first four K-steps of
// O += P V run while the second chunk is still being exponentiated.
template <int POLY, bool SPLIT>
__global__ void __maxnreg__(96)
attention_tc_persistent_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tm_p,
                               const __grid_constant__ CUtensorMap tm_c, int S, int heads, int n_qb, int n_items, int Np,
                               int has_cond, bf16* __restrict__ out_c) {
  extern __shared__ uint8_t smem_raw[];
  if ((ptx::smem_u32(smem_raw) & 1023u) != 0) __trap();  // swizzled TMA tiles need 1024-byte alignment
  uint8_t* sQ = smem_raw;                                  // [2][TILE_BYTES]
  uint8_t* sR = sQ + 2 * TILE_BYTES;                       // [RING_P][TILE_BYTES]: K_0 V_0 ... K_{n-1} V_{n-1} [O staging]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sR + RING_P * TILE_BYTES);
  uint64_t* q_full = bars;                   // 1
  uint64_t* q_empty = bars + 1;              // 1
  uint64_t* r_full = bars + 2;               // RING_P
  uint64_t* r_empty = r_full + RING_P;       // RING_P
  uint64_t* s_full = r_empty + RING_P;       // [tile]
  uint64_t* p_full = s_full + 2;             // [tile] (8 arrivals: one per softmax warp)
  uint64_t* o_full = p_full + 2;             // [tile]
  uint64_t* o_empty = o_full + 2;            // [tile] (8 arrivals)
  uint64_t* p_half = o_empty + 2;            // [tile] (8 arrivals; SPLIT: the first 64 keys of P are in TMEM)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_half + 2);
  float* xch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d = heads * HD;
  const int n_tiles = (S + BKV - 1) / BKV;
  long long* trace = blockIdx.x == 0 && lane == 0 ? g_trace : nullptr;
  if (warp == 0) LC_TRACE(3, 0, 0);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm);
    ptx::prefetch_tmap(&tm_p);
    if (has_cond) ptx::prefetch_tmap(&tm_c);
    ptx::mbar_init(q_full, 1);
    ptx::mbar_init(q_empty, 1);
    for (int i = 0; i < RING_P; ++i) {
      ptx::mbar_init(&r_full[i], 1);
      ptx::mbar_init(&r_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&s_full[i], 1);
      ptx::mbar_init(&p_full[i], 8);
      ptx::mbar_init(&o_full[i], 1);
      ptx::mbar_init(&o_empty[i], 8);
      ptx::mbar_init(&p_half[i], 8);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<TMEM_COLS>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_grid_sync();  // q|k|v are produced by the preceding kernel: nothing global is touched above this line
  if (warp == 0) LC_TRACE(3, 0, 1);

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    int st = 0, k = 0;
    uint32_t ph = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++k) {
      const int qb = item % n_qb, bh = item / n_qb;
      const int h = bh % heads, b = bh / heads;
      const int row_base = b * S, q0 = qb * (2 * BQ);
      if (k > 0) ptx::mbar_wait(q_empty, (k - 1) & 1);  // every S = Q K^T of the previous item has completed
      ptx::mbar_expect_tx(q_full, 2 * TILE_BYTES);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        ptx::tma_load_2d(sQ + t * TILE_BYTES, &tm, q_full, h * HD, row_base + q0 + t * BQ);
        ptx::tma_load_2d(sQ + t * TILE_BYTES + SUB_BYTES, &tm, q_full, h * HD + 64, row_base + q0 + t * BQ);
      }
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
          if (++st == RING_P) { st = 0; ph ^= 1; }
        }
      }
      // the two slots behind V_{n-1} are handed to the softmax warps as the item's output staging tiles (A, B)
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        ptx::mbar_wait(&r_empty[st], ph ^ 1);
        ptx::mbar_arrive(&r_full[st]);
        if (++st == RING_P) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp walks the loop converged (waits, counters, addresses are warp-uniform) and ONE elected lane issues:
    // inside an elect.sync region the compiler keeps descriptors in uniform registers, so a K-step costs ~4 SASS
    // instructions.  With the loop inside `if (lane == 0)` every MMA was preceded by a ~15-instruction
    // ELECT / R2UR.BROADCAST sequence and the issue thread, not the tensor pipe, set the pace (81-94 cycles per M128
    // MMA against the 68 the pipe needs).
    constexpr uint32_t idesc_qk = ptx::make_idesc_bf16(128, BKV, 0, 0);  // A = Q (K-major), B = K (K-major)
    constexpr uint32_t idesc_pv = ptx::make_idesc_bf16(128, HD, 0, 1);   // A = P (TMEM), B = V (MN-major)
    constexpr uint32_t desc_hi = ptx::smem_desc_hi(1024);
    const uint32_t q_addr = ptx::smem_u32(sQ), r_addr = ptx::smem_u32(sR);
    auto issue_qk = [&](int t, int slot) {
      const uint32_t a_lo = ptx::smem_desc_lo(q_addr + t * TILE_BYTES, 16);
      const uint32_t b_lo = ptx::smem_desc_lo(r_addr + slot * TILE_BYTES, 16);
      const uint32_t d_tmem = tmem_base + COL_S + static_cast<uint32_t>(t * 128);
      if (ptx::elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {  // 128 head dims = 2 sub-tiles x 4 K-steps of 16
          const uint32_t off = ((kk >> 2) * SUB_BYTES + (kk & 3) * 32) >> 4;
          ptx::umma_f16_lh(d_tmem, a_lo + off, b_lo + off, desc_hi, idesc_qk, kk != 0 ? 1u : 0u);
        }
        ptx::umma_commit(&s_full[t]);
      }
      __syncwarp();
    };
    auto issue_pv = [&](int t, int slot, bool first, int kk0, int kk1) {
      const uint32_t b_lo = ptx::smem_desc_lo(r_addr + slot * TILE_BYTES, SUB_BYTES);
      const uint32_t d_tmem = tmem_base + COL_O + static_cast<uint32_t>(t * 128);
      const uint32_t a_tmem = tmem_base + COL_S + static_cast<uint32_t>(t * 128);
      if (ptx::elect_one()) {
#pragma unroll
        for (int kk = kk0; kk < kk1; ++kk)  // 128 keys = 8 K-steps of 16 (16 key rows x 128 B = 2 KB per sub-tile)
          ptx::umma_f16_ts_lh(d_tmem, a_tmem + kk * 8, b_lo + ((kk * 2048) >> 4), desc_hi, idesc_pv,
                              (first && kk == 0) ? 0u : 1u);
      }
      __syncwarp();
    };
    auto commit = [&](uint64_t* bar) {
      if (ptx::elect_one()) ptx::umma_commit(bar);
      __syncwarp();
    };
    int st = 0, k = 0;
    uint32_t ph = 0, sc = 0;  // sc: running (item, step) count = phase counter of s_full / p_full
    auto advance = [&]() { if (++st == RING_P) { st = 0; ph ^= 1; } };
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++k) {
      ptx::mbar_wait(q_full, k & 1);
      ptx::mbar_wait(&r_full[st], ph);
      ptx::tc_fence_after();
      // the S columns are free: the previous item's last P V were issued before (the tensor pipe runs in order)
      issue_qk(0, st);
      issue_qk(1, st);
      commit(&r_empty[st]);
      if (n_tiles == 1) commit(q_empty);
      advance();
      for (int j = 0; j < n_tiles; ++j, ++sc) {
        const int v_slot = st;
        ptx::mbar_wait(&r_full[v_slot], ph);
        advance();
        const int k_slot = st;  // K_{j+1}
        const uint32_t k_ph = ph;
        const bool more = j + 1 < n_tiles;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (SPLIT) {
            ptx::mbar_wait(&p_half[t], sc & 1);
            if (j == 0 && k > 0) ptx::mbar_wait(&o_empty[t], (k - 1) & 1);  // the epilogue has drained O_t
            ptx::tc_fence_after();
            if (k == 0) LC_TRACE(0, j, 2 * t);
            issue_pv(t, v_slot, j == 0, 0, 4);
            ptx::mbar_wait(&p_full[t], sc & 1);
            ptx::tc_fence_after();
            issue_pv(t, v_slot, false, 4, 8);
          } else {
            ptx::mbar_wait(&p_full[t], sc & 1);
            if (j == 0 && k > 0) ptx::mbar_wait(&o_empty[t], (k - 1) & 1);  // the epilogue has drained O_t
            ptx::tc_fence_after();
            if (k == 0) LC_TRACE(0, j, 2 * t);
            issue_pv(t, v_slot, j == 0, 0, 8);
          }
          if (!more) commit(&o_full[t]);
          if (t == 1) commit(&r_empty[v_slot]);
          if (more) {
            if (t == 0) {
              ptx::mbar_wait(&r_full[k_slot], k_ph);
              ptx::tc_fence_after();
            }
            issue_qk(t, k_slot);
            if (t == 1) {
              commit(&r_empty[k_slot]);
              if (j + 2 == n_tiles) commit(q_empty);  // the item's last S tiles: Q may be replaced
            }
          }
          if (k == 0) LC_TRACE(0, j, 2 * t + 1);
        }
        if (more) advance();
      }
      advance();  // the two slots behind V_{n-1} are the item's output staging tiles (owned by the softmax warps)
      advance();
    }
  } else if (warp >= 2) {
    // ===================== softmax + epilogue =====================
    const int idx = warp - 2;
    const int t = idx >> 3;          // 0: tile A (warps 2-9), 1: tile B (warps 10-17)
    const int hh = (idx >> 2) & 1;   // key half of the row this thread exponentiates
    const int quarter = warp & 3;    // TMEM lane quarter this warp may access (= warp id % 4)
    if (idx & 7) trace = nullptr;    // one traced warp per tile
    const int r = quarter * 32 + lane;
    const uint32_t pair_bar = 1 + t * 4 + quarter;  // named barrier of the two warps sharing these 32 rows
    const uint32_t tile_bar = 9 + t;                // named barrier of the tile's 8 warps
    const bool storer = (idx & 7) == 0 && lane == 0;  // issues the tile's TMA stores
    int pending_slot = -1;                            // staging slot whose stores may still be reading it
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const float scale_log2 = 0.08838834764831845f * 1.4426950408889634f;
    const uint32_t s_addr = tmem_base + lane_addr + COL_S + static_cast<uint32_t>(t * 128);
    const uint32_t o_addr = tmem_base + lane_addr + COL_O + static_cast<uint32_t>(t * 128 + hh * 64);
    uint32_t sc = 0;
    int k = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++k) {
      const int qb = item % n_qb, bh = item / n_qb;
      const int h = bh % heads, b = bh / heads;
      const int q0 = qb * (2 * BQ);
      long long* tr = k == 0 ? trace : nullptr;
      float m_used = -INFINITY, l = 0.f;
      for (int j = 0; j < n_tiles; ++j, ++sc) {
        if (tr != nullptr) tr[((1 + t) * 64 + j) * 4 + 0] = clock64();
        ptx::mbar_wait(&s_full[t], sc & 1);
        ptx::tc_fence_after();
        if (tr != nullptr) tr[((1 + t) * 64 + j) * 4 + 1] = clock64();
        if (trace != nullptr && k == 1 && j == 0) trace[(3 * 64 + 1 + t) * 4 + 3] = clock64();  // next item's first S ready
        if (pending_slot >= 0) {  // (storer only) the previous item's stores have drained the staging tile by now
          ptx::bulk_wait_read();
          ptx::mbar_arrive(&r_empty[pending_slot]);  // back to the producer: the slot carries K/V again
          pending_slot = -1;
        }
        const int n_valid = S - j * BKV;  // keys >= n_valid are padding (only ever true for the last tile)
        // key offset of sreg[c][0]: one 64-key half, or (SPLIT) 32 keys of each 64-key chunk
        const int key0 = SPLIT ? hh * 32 : hh * 64, key1 = SPLIT ? 64 + hh * 32 : hh * 64 + 32;
        uint32_t sreg[2][32];
        ptx::tmem_ld32(s_addr + key0, sreg[0]);
        ptx::tmem_ld32(s_addr + key1, sreg[1]);
        ptx::tmem_ld_wait();
        if (n_valid < BKV) {
#pragma unroll
          for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if ((c == 0 ? key0 : key1) + i >= n_valid) sreg[c][i] = 0xff800000u;  // -inf
        }
        float neg_m = -m_used;
        float ps[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t pk[32];
        auto exps = [&](auto lo) {  // exponentials + bf16 packing of elements [lo, lo + 32) of this thread's 64
          constexpr int LO = decltype(lo)::value;
#pragma unroll
          for (int i = LO; i < LO + 32; i += 2) {
            const float x0 = fmaf(__uint_as_float(sreg[i >> 5][i & 31]), scale_log2, neg_m);
            const float x1 = fmaf(__uint_as_float(sreg[i >> 5][(i & 31) + 1]), scale_log2, neg_m);
            const float p0 = (POLY > 0 && (i % POLY) == POLY - 1) ? ptx::ex2_poly(x0) : ptx::ex2_approx(x0);
            const float p1 = (POLY > 0 && ((i + 1) % POLY) == POLY - 1) ? ptx::ex2_poly(x1) : ptx::ex2_approx(x1);
            ps[(i >> 1) & 3] += p0 + p1;
            __nv_bfloat162 pb = __floats2bfloat162_rn(p0, p1);
            pk[i >> 1] = *reinterpret_cast<uint32_t*>(&pb);
          }
        };
        // (Measured and dropped: exponentiating the first chunk speculatively against the previous running maximum
        // while the row maximum is still being exchanged — 1009 vs 1135 TFLOP/s, the extra live registers spill.)
        float mxs[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i) mxs[i & 3] = fmaxf(mxs[i & 3], __uint_as_float(sreg[c][i]));
        float mx = fmaxf(fmaxf(mxs[0], mxs[1]), fmaxf(mxs[2], mxs[3]));
        xch[(t * 2 + hh) * 128 + r] = mx;
        asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
        mx = fmaxf(mx, xch[(t * 2 + (1 - hh)) * 128 + r]);
        if (tr != nullptr) tr[((1 + t) * 64 + j) * 4 + 2] = clock64();
        const float m_new = fmaxf(m_used, mx * scale_log2);
        const bool need = __any_sync(0xffffffffu, m_new > m_used + RESCALE_THRESHOLD);
        float alpha = 1.f;
        if (need) {
          alpha = (m_used == -INFINITY) ? 0.f : exp2f(m_used - m_new);
          m_used = m_new;
          neg_m = -m_used;
        }
        exps(std::integral_constant<int, 0>{});
        if (SPLIT && need && j > 0) {  // rare: the O rescale must precede the FIRST chunk's hand-over
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t o[16];
            ptx::tmem_ld16(o_addr + c * 16, o);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            ptx::tmem_st16r(o_addr + c * 16, o);
          }
        }
        if (SPLIT) {  // P of keys [0, 64) (this thread: P columns [16 hh, 16 hh + 16)) -> tensor pipe
          ptx::tmem_st16<0>(s_addr + hh * 16, pk);
          ptx::tmem_st_wait();
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&p_half[t]);
        }
        exps(std::integral_constant<int, 32>{});
        if (!SPLIT && need && j > 0) {  // rare: rescale this thread's 64 columns of O (quiescent: PV_t(j-1) has completed)
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t o[16];
            ptx::tmem_ld16(o_addr + c * 16, o);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            ptx::tmem_st16r(o_addr + c * 16, o);
          }
        }
        if (SPLIT) ptx::tmem_st16<16>(s_addr + 32 + hh * 16, pk);  // P of keys [64, 128)
        else ptx::tmem_st32(s_addr + hh * 32, pk);
        l = l * alpha + ((ps[0] + ps[1]) + (ps[2] + ps[3]));
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&p_full[t]);
        if (tr != nullptr) tr[((1 + t) * 64 + j) * 4 + 3] = clock64();
      }
      // ---- epilogue of the item: O_t / l -> bf16 -> swizzled staging tile -> TMA tensor stores
      ptx::mbar_wait(&o_full[t], k & 1);
      ptx::tc_fence_after();
      if (tr != nullptr) tr[(3 * 64 + 3 + t) * 4 + 0] = clock64();  // O_t complete
      if (tr != nullptr) tr[(3 * 64 + 0) * 4 + 2] = clock64();
      xch[(t * 2 + hh) * 128 + r] = l;
      asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
      const float inv = 1.0f / (l + xch[(t * 2 + (1 - hh)) * 128 + r]);
      asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");  // both have read before the slots are reused
      // staging tile of tile t = ring position k * (2 n + 2) + 2 n + t (behind the item's last V)
      const int pos = k * (2 * n_tiles + 2) + 2 * n_tiles + t;
      const int stg_slot = pos % RING_P;
      ptx::mbar_wait(&r_full[stg_slot], (pos / RING_P) & 1);
      if (tr != nullptr) tr[(3 * 64 + 1 + t) * 4 + 0] = clock64();
      uint8_t* sO = sR + stg_slot * TILE_BYTES;
      uint8_t* srow = sO + hh * SUB_BYTES + r * 128;
      // The one tile per (sample, head) that straddles the pred | cond boundary: its pred rows go through the TMA store
      // (clipped at Np), its cond rows would need a negative start coordinate in the cond map, so the few threads that
      // own them store their 128 bytes directly.
      const int tok0 = q0 + t * BQ, tok = tok0 + r;
      bf16* direct = nullptr;
      if (has_cond && tok0 < Np && tok >= Np && tok < S)
        direct = out_c + (static_cast<long long>(b) * (S - Np) + (tok - Np)) * d + h * HD + hh * 64;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t o[32];
        ptx::tmem_ld32(o_addr + c * 32, o);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 pkt;
          __nv_bfloat162 t0 = __floats2bfloat162_rn(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv);
          __nv_bfloat162 t1 = __floats2bfloat162_rn(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv);
          __nv_bfloat162 t2 = __floats2bfloat162_rn(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv);
          __nv_bfloat162 t3 = __floats2bfloat162_rn(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv);
          pkt.x = *reinterpret_cast<uint32_t*>(&t0);
          pkt.y = *reinterpret_cast<uint32_t*>(&t1);
          pkt.z = *reinterpret_cast<uint32_t*>(&t2);
          pkt.w = *reinterpret_cast<uint32_t*>(&t3);
          const int chunk = c * 4 + (i >> 3);  // 16-byte chunk of the 128-byte row, 128B-swizzled like a TMA tile
          *reinterpret_cast<uint4*>(srow + ((chunk ^ (r & 7)) << 4)) = pkt;
          if (direct != nullptr) *reinterpret_cast<uint4*>(direct + c * 32 + i) = pkt;
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&o_empty[t]);  // O_t may be overwritten by the next item's first P V
      if (tr != nullptr) tr[(3 * 64 + 1 + t) * 4 + 1] = clock64();
      ptx::fence_proxy_async();  // staging writes -> visible to the TMA (async proxy)
      asm volatile("bar.sync %0, 256;" ::"r"(tile_bar) : "memory");
      if (tr != nullptr) tr[(3 * 64 + 1 + t) * 4 + 2] = clock64();
      if (storer) {
        // rows beyond the stream's length are clipped by the tensor map (coordinates are never negative)
        if (tok0 < Np) {
          ptx::tma_store_3d(&tm_p, sO, h * HD, tok0, b);
          ptx::tma_store_3d(&tm_p, sO + SUB_BYTES, h * HD + 64, tok0, b);
        } else if (has_cond && tok0 < S) {
          ptx::tma_store_3d(&tm_c, sO, h * HD, tok0 - Np, b);
          ptx::tma_store_3d(&tm_c, sO + SUB_BYTES, h * HD + 64, tok0 - Np, b);
        }
        ptx::bulk_commit();
        pending_slot = stg_slot;
        if (tr != nullptr) tr[(3 * 64 + 0) * 4 + 3] = clock64();
      }
    }
    if (pending_slot >= 0) {  // last item: the stores must have read the tile before the CTA's shared memory goes away
      ptx::bulk_wait_read();
      ptx::mbar_arrive(&r_empty[pending_slot]);
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
  const int d = heads * HD;
  CUtensorMap tm;
  LC_TRY(make_tmap_2d_bf16(&tm, qkv, static_cast<uint64_t>(3) * d, static_cast<uint64_t>(B) * S,
                           static_cast<uint64_t>(3) * d * 2, 64, 128));
  const double flops = 4.0 * B * heads * static_cast<double>(S) * S * HD, bytes = 8.0 * B * S * d;  // q, k, v in + o out
  static const int variant = [] { const char* e = getenv("LADCAST_B200_ATTN"); return (e != nullptr && e[0] == '6') ? 6 : 7; }();
  if (variant == 6) {  // one work item per CTA (round 1)
    static PerDevice<bool> attr_set;
    if (!attr_set.here()) {
      LC_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
      attr_set.here() = true;
    }
    dim3 grid(ceil_div(S, 2 * BQ), heads, B);
    prof_begin(PROF_ATTN, s);
    LC_CHECK_CUDA(launch_kernel(attention_tc_kernel, grid, NUM_THREADS, SMEM_BYTES, s, tm, S, heads, out_p, Np, out_c));
    prof_end(PROF_ATTN, flops, s, bytes);
    LC_LAUNCH_CHECK();
    return 0;
  }
  // persistent: one CTA per SM over (sample, head, query block) items; outputs through 3-D TMA store maps
  // {d, tokens of the stream, samples} so that the hardware clips rows past the stream's length
  // LADCAST_B200_ATTN_POLY = n: every n-th exponential on the FMA / ALU pipes (0 = all on the MUFU)
  static const int poly = [] { const char* e = getenv("LADCAST_B200_ATTN_POLY"); return e != nullptr ? atoi(e) : ATTN_POLY_DEFAULT; }();
  static const int split = [] { const char* e = getenv("LADCAST_B200_ATTN_SPLIT"); return e != nullptr ? atoi(e) : ATTN_SPLIT_DEFAULT; }();
  auto kern = !split ? attention_tc_persistent_kernel<0, false>
              : poly == 3 ? attention_tc_persistent_kernel<3, true> : attention_tc_persistent_kernel<0, true>;
  static PerDevice<bool> attr_set_p;
  if (!attr_set_p.here()) {
    LC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES_P));
    attr_set_p.here() = true;
  }
  if (Np == 0) {  // every token belongs to the second stream (context refiner): treat it as the only stream
    out_p = out_c;
    out_c = nullptr;
    Np = S;
  }
  LC_REQUIRE(out_p != nullptr && Np >= 1 && Np <= S && (out_c != nullptr || Np == S), "attention: bad pred / cond split");
  CUtensorMap tm_p, tm_c;
  {
    uint64_t dims[3] = {static_cast<uint64_t>(d), static_cast<uint64_t>(Np), static_cast<uint64_t>(B)};
    uint64_t strides[2] = {static_cast<uint64_t>(d) * 2, static_cast<uint64_t>(Np) * d * 2};
    uint32_t box[3] = {64, 128, 1};
    LC_TRY(make_tmap_bf16(&tm_p, out_p, 3, dims, strides, box));
  }
  const int has_cond = (Np < S) ? 1 : 0;
  if (has_cond) {
    uint64_t dims[3] = {static_cast<uint64_t>(d), static_cast<uint64_t>(S - Np), static_cast<uint64_t>(B)};
    uint64_t strides[2] = {static_cast<uint64_t>(d) * 2, static_cast<uint64_t>(S - Np) * d * 2};
    uint32_t box[3] = {64, 128, 1};
    LC_TRY(make_tmap_bf16(&tm_c, out_c, 3, dims, strides, box));
  } else {
    tm_c = tm_p;
  }
  const int n_qb = ceil_div(S, 2 * BQ);
  const int n_items = n_qb * heads * B;
  const int grid = n_items < num_sms() ? n_items : num_sms();
  prof_begin(PROF_ATTN, s);
  LC_CHECK_CUDA(launch_kernel(kern, dim3(grid), NUM_THREADS_P, SMEM_BYTES_P, s, tm, tm_p, tm_c, S, heads,
                              n_qb, n_items, Np, has_cond, out_c));
  prof_end(PROF_ATTN, flops, s, bytes);
  LC_LAUNCH_CHECK();
  return 0;
}

}  // namespace lc
