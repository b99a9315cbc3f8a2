// tcgen05 / TMEM / TMA GEMM for sm_100a:  C[M,N] = A[M,K] * W[N,K]^T with fused epilogues (common.cuh::EpiParams).
//
//   * persistent: one CTA per SM loops over 128 x BN output tiles (grouped rasterisation for L2 reuse of W);
//   * warp 0 (one lane)  : TMA producer, 4-6 stage smem ring of {A 128x64, W BNx64} bf16 tiles, 128B swizzle;
//   * warp 1 (one lane)  : issues tcgen05.mma (M=128, N=BN, K=16), accumulators in TMEM, double buffered;
//   * warp 2             : TMEM allocation / deallocation;
//   * warps 4-7          : epilogue, tcgen05.ld 32 lanes x 32 columns -> registers -> bias/act/gate -> global.
//
// A may come from two K-segments (two tensor maps) so that the single-stream block's proj_out reads
// [attention | mlp] without a concat (reference: LaDCast_3D_model.py:460-461).
// The same kernel serves as an implicit-GEMM 3x3 sphere convolution (A tiles fetched from a padded NHWC tensor
// with a 4-D tensor map at tap-shifted coordinates; reference: models/sphere_conv.py:138-192).
#include <cstdlib>

#include "common.cuh"
#include "gemm_tc.h"
#include <cuda_fp16.h>
#include "ptx.cuh"
#include "tmap.h"

namespace lc {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr int EPI_WARP0 = 4;
constexpr int NUM_EPI_WARPS = 8;  // two per TMEM lane quarter, each owning half of the tile's columns
constexpr int NUM_THREADS = (EPI_WARP0 + NUM_EPI_WARPS) * 32;
constexpr int EPI_SMEM = NUM_EPI_WARPS * 32 * 33 * 4;  // per-warp padded 32x32 transpose tiles

template <int BN>
struct Cfg {
  static constexpr int STAGE_A = BM * BK * 2;
  static constexpr int STAGE_B = BN * BK * 2;
  static constexpr int STAGE = STAGE_A + STAGE_B;
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int TMEM_COLS = 2 * BN;  // two accumulator stages
  static constexpr int SMEM = STAGES * STAGE + EPI_SMEM + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct TileSched {
  int num_m, num_n, group_m;
  __device__ __forceinline__ void decode(int tile, int& m_blk, int& n_blk) const {
    const int per_group = group_m * num_n;
    const int g = tile / per_group;
    const int first_m = g * group_m;
    const int gm = min(group_m, num_m - first_m);
    const int within = tile - g * per_group;
    m_blk = first_m + within % gm;
    n_blk = within / gm;
  }
};

// Fast activations for the bf16 path (MUFU tanh / ex2): the epilogue runs on 8 warps next to a saturated tensor
// pipe, so it must stay a few instructions per element.  (The FP32 validation path uses the precise versions.)
template <int ACT>
__device__ __forceinline__ float act_fast(float v) {
  if (ACT == ACT_GELU_TANH) {
    const float u = 0.7978845608028654f * (v + 0.044715f * v * v * v);
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
    return 0.5f * v * (1.0f + t);
  }
  if (ACT == ACT_SILU) {
    // x * sigmoid(x) = 0.5 x (1 + tanh(x / 2)): ONE MUFU op per element (ex2 + rcp would be two; the 1x1 C -> 8C
    // projections of the DC-AE write 8064 SiLU outputs per pixel and their epilogue was the longer leg)
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * v));
    return 0.5f * v * (1.0f + t);
  }
  if (ACT == ACT_RELU) return fmaxf(v, 0.0f);
  return v;
}

// Optional in-kernel timeline of pair 0 for tuning (tools/gemm_trace.py): clock64 stamps, [role][tile][4].
__device__ long long* g_gemm_trace = nullptr;

enum EpiKind { K_STORE_BF16 = 0, K_STORE_F32 = 1, K_GATED = 2, K_RESID = 3, K_UNPATCH = 4, K_QKV = 5, K_NORM = 6 };

// One 32x32 accumulator block (lane = row, r[j] = column n0+j).  K_UNPATCH keeps this layout (consecutive rows are
// contiguous in the channel-major output); every other kind transposes the block through a padded smem tile so
// that lanes run along the columns and each global access is a contiguous 64-128 B row segment.  Kind and
// activation are template parameters: the row loops are branch-free.
template <int KIND, int ACT>
__device__ __forceinline__ void epilogue_block(const EpiParams& ep, const uint32_t (&r)[32], float* tbuf, int lane,
                                               int row_mine, long long orow_mine, int sample_mine, int n0, int M, int N) {
  if (KIND == K_UNPATCH) {
    if (row_mine >= M) return;
    const int r_in = row_mine - sample_mine * ep.rows_per_sample;
    long long chs;
    float* op = reinterpret_cast<float*>(ep.out) + epi_unpatch_base(ep, sample_mine, chs) + n0 * chs + r_in;
    const bool has_bias = ep.bias != nullptr, has_scale = ep.ch_scale != nullptr;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (n0 + j < ep.n_valid) {
        float o = __uint_as_float(r[j]);
        if (has_bias) o += __ldg(ep.bias + n0 + j);
        o = act_fast<ACT>(o);
        if (has_scale) o = o * __ldg(ep.ch_scale + n0 + j) + __ldg(ep.ch_shift + n0 + j);
        op[j * chs] = o;
      }
    }
    return;
  }
  if (KIND == K_STORE_BF16 && n0 + 32 <= N && (ep.ldo & 7) == 0) {
    // bf16 output, full chunk: each thread owns 32 consecutive columns of its row = 64 contiguous bytes.  Four
    // 16-byte stores per thread need ~4x fewer instructions than the transposed path and the epilogue of the
    // K=1536 projections is instruction-bound; L2 merges the half-sector writes before they reach HBM.
    if (row_mine >= M) return;
    float v[32];
    if (ep.bias != nullptr) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + n0 + j));
        v[j] = __uint_as_float(r[j]) + b4.x; v[j + 1] = __uint_as_float(r[j + 1]) + b4.y;
        v[j + 2] = __uint_as_float(r[j + 2]) + b4.z; v[j + 3] = __uint_as_float(r[j + 3]) + b4.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
    }
    uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(ep.out) + orow_mine * ep.ldo + n0);
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      __nv_bfloat162 t0 = __floats2bfloat162_rn(act_fast<ACT>(v[j]), act_fast<ACT>(v[j + 1]));
      __nv_bfloat162 t1 = __floats2bfloat162_rn(act_fast<ACT>(v[j + 2]), act_fast<ACT>(v[j + 3]));
      __nv_bfloat162 t2 = __floats2bfloat162_rn(act_fast<ACT>(v[j + 4]), act_fast<ACT>(v[j + 5]));
      __nv_bfloat162 t3 = __floats2bfloat162_rn(act_fast<ACT>(v[j + 6]), act_fast<ACT>(v[j + 7]));
      uint4 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&t0);
      pk.y = *reinterpret_cast<uint32_t*>(&t1);
      pk.z = *reinterpret_cast<uint32_t*>(&t2);
      pk.w = *reinterpret_cast<uint32_t*>(&t3);
      op[j >> 3] = pk;
    }
    return;
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) tbuf[lane * 33 + j] = __uint_as_float(r[j]);
  __syncwarp();
  if (KIND == K_STORE_BF16) {
    // half-warps cover two rows per instruction, each lane packs two adjacent columns (64 B per row)
    const int cc = (lane & 15) * 2;
    const int col = n0 + cc;
    const float b0 = (ep.bias != nullptr && col < N) ? __ldg(ep.bias + col) : 0.f;
    const float b1 = (ep.bias != nullptr && col + 1 < N) ? __ldg(ep.bias + col + 1) : 0.f;
    bf16* outp = reinterpret_cast<bf16*>(ep.out) + col;
    const long long ldo = ep.ldo;
    const bool pair_ok = col + 1 < N, one_ok = col < N;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int rr = 2 * i + (lane >> 4);
      const int row = __shfl_sync(0xffffffffu, row_mine, rr);
      const long long orow = __shfl_sync(0xffffffffu, orow_mine, rr);
      const float v0 = act_fast<ACT>(tbuf[rr * 33 + cc] + b0);
      const float v1 = act_fast<ACT>(tbuf[rr * 33 + cc + 1] + b1);
      if (row < M) {
        if (pair_ok) *reinterpret_cast<__nv_bfloat162*>(outp + orow * ldo) = __floats2bfloat162_rn(v0, v1);
        else if (one_ok) outp[orow * ldo] = __float2bfloat16_rn(v0);
      }
    }
  } else if (n0 + 32 <= N && (ep.ldo & 3) == 0 && (KIND != K_GATED || (ep.gate_stride & 3) == 0) &&
             (KIND != K_RESID || (ep.ldr & 3) == 0)) {
    // fp32 output, full chunk: 8 lanes x float4 cover one 128-B row segment, a warp instruction covers 4 rows, and
    // all 8 row groups' residual / gate loads are issued before the first use: 4 KB of residual in flight per warp.
    // (With one 4-B column per lane the read-modify-write epilogue of the K=1536 out-projections was bound by
    // DRAM latency x 1 KB in flight per warp: 1.6 TB/s.)
    const int c4 = (lane & 7) * 4, rsub = lane >> 3;
    const int col = n0 + c4;
    float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ep.bias != nullptr) bias4 = __ldg(reinterpret_cast<const float4*>(ep.bias + col));
    float* outp = reinterpret_cast<float*>(ep.out) + col;
    const long long ldo = ep.ldo;
    float4 prev[8], gq[8];
    long long orows[8];
    bool ok[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const int rr = g * 4 + rsub;
      ok[g] = __shfl_sync(0xffffffffu, row_mine, rr) < M;
      orows[g] = __shfl_sync(0xffffffffu, orow_mine, rr);
      prev[g] = make_float4(0.f, 0.f, 0.f, 0.f);
      gq[g] = make_float4(1.f, 1.f, 1.f, 1.f);
      if (KIND == K_GATED) {
        const int smp = __shfl_sync(0xffffffffu, sample_mine, rr);
        if (ok[g]) {
          prev[g] = *reinterpret_cast<const float4*>(outp + orows[g] * ldo);
          if (ep.gate != nullptr)
            gq[g] = __ldg(reinterpret_cast<const float4*>(ep.gate + static_cast<long long>(smp) * ep.gate_stride + col));
        }
      } else if (KIND == K_RESID) {
        if (ok[g]) prev[g] = *reinterpret_cast<const float4*>(ep.resid + orows[g] * ep.ldr + col);
      }
    }
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const float* tb = tbuf + (g * 4 + rsub) * 33 + c4;
      const float t0 = tb[0] + bias4.x, t1 = tb[1] + bias4.y, t2 = tb[2] + bias4.z, t3 = tb[3] + bias4.w;
      float4 o;
      if (KIND == K_GATED) {
        o = make_float4(fmaf(gq[g].x, t0, prev[g].x), fmaf(gq[g].y, t1, prev[g].y), fmaf(gq[g].z, t2, prev[g].z),
                        fmaf(gq[g].w, t3, prev[g].w));
      } else {
        o = make_float4(act_fast<ACT>(t0) + prev[g].x, act_fast<ACT>(t1) + prev[g].y, act_fast<ACT>(t2) + prev[g].z,
                        act_fast<ACT>(t3) + prev[g].w);
      }
      if (ok[g]) *reinterpret_cast<float4*>(outp + orows[g] * ldo) = o;
    }
  } else {
    const int col = n0 + lane;
    const bool col_ok = col < N;
    const float bias = (ep.bias != nullptr && col_ok) ? __ldg(ep.bias + col) : 0.f;
    float* outp = reinterpret_cast<float*>(ep.out) + col;
    const long long ldo = ep.ldo;
#pragma unroll
    for (int r0 = 0; r0 < 32; r0 += 8) {
      // issue the 8 residual/gate loads of this row group first (memory-level parallelism), then combine + store
      float prev[8], gq[8];
      bool ok[8];
      long long orows[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        ok[q] = (__shfl_sync(0xffffffffu, row_mine, r0 + q) < M) && col_ok;
        orows[q] = __shfl_sync(0xffffffffu, orow_mine, r0 + q);
        prev[q] = 0.f;
        gq[q] = 1.f;
        if (KIND == K_GATED) {
          const int smp = __shfl_sync(0xffffffffu, sample_mine, r0 + q);
          if (ok[q]) {
            prev[q] = outp[orows[q] * ldo];
            if (ep.gate != nullptr) gq[q] = __ldg(ep.gate + static_cast<long long>(smp) * ep.gate_stride + col);
          }
        } else if (KIND == K_RESID) {
          if (ok[q]) prev[q] = ep.resid[orows[q] * ep.ldr + col];
        }
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float t = tbuf[(r0 + q) * 33 + lane] + bias;
        if (ok[q]) {
          if (KIND == K_GATED) outp[orows[q] * ldo] = fmaf(gq[q], t, prev[q]);
          else outp[orows[q] * ldo] = act_fast<ACT>(t) + prev[q];
        }
      }
    }
  }
  __syncwarp();
}

// EPI_NORM_RESID chunk: the 32x32 block is transposed through shared memory so that 8 lanes x float4 cover a 128-byte
// row segment; the row's rstd travels with the row (shuffle), RMSNorm weight / bias with the column.  x += y in place
// (fp32), the new x is also stored as bf16 at the remapped row (next conv's padded input / next GEMM's operand).
__device__ __forceinline__ void epilogue_block_norm(const EpiParams& ep, const uint32_t (&r)[32], float* tbuf, int lane,
                                                    int row_mine, long long orow_mine, float rstd_mine, int n0, int M,
                                                    int N) {
#pragma unroll
  for (int j = 0; j < 32; ++j) tbuf[lane * 33 + j] = __uint_as_float(r[j]);
  __syncwarp();
  const int c4 = (lane & 7) * 4, rsub = lane >> 3;
  const int col = n0 + c4;
  const bool col_ok = col < N;  // N % 4 == 0 (checked on the host)
  float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f), b4 = w4;
  if (col_ok) {
    w4 = __ldg(reinterpret_cast<const float4*>(ep.norm_w + col));
    b4 = __ldg(reinterpret_cast<const float4*>(ep.norm_b + col));
  }
  float4 prev[8];
  long long orows[8];
  int rows[8];
  float rs[8];
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const int rr = g * 4 + rsub;
    rows[g] = __shfl_sync(0xffffffffu, row_mine, rr);
    orows[g] = __shfl_sync(0xffffffffu, orow_mine, rr);
    rs[g] = __shfl_sync(0xffffffffu, rstd_mine, rr);
    prev[g] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rows[g] < M && col_ok && ep.xres != nullptr)
      prev[g] = *reinterpret_cast<const float4*>(ep.xres + static_cast<long long>(rows[g]) * ep.ldr + col);
  }
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const float* tb = tbuf + (g * 4 + rsub) * 33 + c4;
    float4 o;
    o.x = fmaf(tb[0] * rs[g], w4.x, b4.x) + prev[g].x;
    o.y = fmaf(tb[1] * rs[g], w4.y, b4.y) + prev[g].y;
    o.z = fmaf(tb[2] * rs[g], w4.z, b4.z) + prev[g].z;
    o.w = fmaf(tb[3] * rs[g], w4.w, b4.w) + prev[g].w;
    if (rows[g] < M && col_ok) {
      if (ep.xres != nullptr) *reinterpret_cast<float4*>(ep.xres + static_cast<long long>(rows[g]) * ep.ldr + col) = o;
      if (ep.out != nullptr) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
        uint2 u;
        u.x = *reinterpret_cast<uint32_t*>(&lo);
        u.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(ep.out) + orows[g] * ep.ldo + col) = u;
      }
    }
  }
  __syncwarp();
}

// qkv projection chunk (32 columns of one head, thread = row): bias, then (q/k heads only) per-head RMSNorm with the
// row's rstd and RoPE on interleaved pairs, bf16, four 16-byte stores.  The rotation factors come from a packed
// [tokens, 64] half2 (cos, sin) table: 64 B per thread and chunk instead of 256 B of fp32 cos + sin.

// Staged bf16 write-back (see drain_bf16_staged): a warp's 32 rows x 64 columns sit in its 4 KB staging tile as
// [row][8 x 16-byte chunk ^ (row & 7)]; lane = (row-in-group rsub = lane / 8, chunk kq = lane % 8): 8 lanes cover one
// contiguous 128-byte row segment, four rows per instruction.
struct StagedRows {
  long long orows[8];
  bool ok[8];
};
__device__ __forceinline__ StagedRows staged_rows(int lane, int row, long long orow, int M) {
  StagedRows sr;
  const int rsub = lane >> 3;
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const int rr = g * 4 + rsub;
    sr.ok[g] = __shfl_sync(0xffffffffu, row, rr) < M;
    sr.orows[g] = __shfl_sync(0xffffffffu, orow, rr);
  }
  return sr;
}
__device__ __forceinline__ void staged_put(uint8_t* buf, int lane, int chunk, const uint4& pk) {
  *reinterpret_cast<uint4*>(buf + lane * 128 + ((chunk ^ (lane & 7)) << 4)) = pk;
}
__device__ __forceinline__ void staged_writeback(const EpiParams& ep, const uint8_t* buf, int lane, int n0_64,
                                                 const StagedRows& sr) {
  const int rsub = lane >> 3, kq = lane & 7;
  __syncwarp();
  bf16* gbase = reinterpret_cast<bf16*>(ep.out) + n0_64 + kq * 8;
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const int rr = g * 4 + rsub;
    const uint4 val = *reinterpret_cast<const uint4*>(buf + rr * 128 + ((kq ^ (rr & 7)) << 4));
    if (sr.ok[g]) *reinterpret_cast<uint4*>(gbase + sr.orows[g] * ep.ldo) = val;
  }
  __syncwarp();
}

// Reference: LaDCast_3D_model.py:92-169 (to_q/k/v, norm_q/k, apply_rotary_emb).
__device__ __forceinline__ void epilogue_qkv_chunk(const EpiParams& ep, const uint32_t (&r)[32], bool row_ok,
                                                   long long orow, int tok, float rstd, int n0, int colh0, bool is_qk,
                                                   const float* nw, uint8_t* stage = nullptr, int lane = 0, int cpar = 0) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ep.bias != nullptr) b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + n0 + j));
    v[j] = __uint_as_float(r[j]) + b4.x; v[j + 1] = __uint_as_float(r[j + 1]) + b4.y;
    v[j + 2] = __uint_as_float(r[j + 2]) + b4.z; v[j + 3] = __uint_as_float(r[j + 3]) + b4.w;
  }
  if (is_qk) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 w4 = __ldg(reinterpret_cast<const float4*>(nw + colh0 + j));
      v[j] *= rstd * w4.x; v[j + 1] *= rstd * w4.y; v[j + 2] *= rstd * w4.z; v[j + 3] *= rstd * w4.w;
    }
    if (ep.rope_cs != nullptr && row_ok) {
      const uint4* cs = reinterpret_cast<const uint4*>(ep.rope_cs + static_cast<long long>(tok) * 64 + (colh0 >> 1));
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 c4 = __ldg(cs + q);
        const uint32_t cw[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&cw[k]));  // (cos, sin)
          const int j = (q * 4 + k) * 2;
          const float a = v[j], b = v[j + 1];
          v[j] = a * f.x - b * f.y;
          v[j + 1] = b * f.x + a * f.y;
        }
      }
    }
  }
  if (!row_ok && stage == nullptr) return;
  uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(ep.out) + orow * ep.ldo + n0);
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    __nv_bfloat162 t0 = __floats2bfloat162_rn(v[j], v[j + 1]), t1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
    __nv_bfloat162 t2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]), t3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
    uint4 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&t0);
    pk.y = *reinterpret_cast<uint32_t*>(&t1);
    pk.z = *reinterpret_cast<uint32_t*>(&t2);
    pk.w = *reinterpret_cast<uint32_t*>(&t3);
    if (stage != nullptr) staged_put(stage, lane, cpar * 4 + (j >> 3), pk);
    else op[j >> 3] = pk;
  }
}

__device__ __forceinline__ int epi_kind(const EpiParams& ep) {
  if (ep.qk_cols > 0) return K_QKV;
  if (ep.mode == EPI_NORM_RESID) return K_NORM;
  if (ep.mode == EPI_GATED_RESID) return K_GATED;
  if (ep.mode == EPI_UNPATCHIFY) return K_UNPATCH;
  if (ep.mode == EPI_RESID_STORE) return K_RESID;  // f32 output only on the tensor-core path
  return ep.out_f32 ? K_STORE_F32 : K_STORE_BF16;
}

// bf16 output, full column slice: bias / activation / packing as in epilogue_block, but the packed rows go through a
// per-warp 4 KB staging tile (32 rows x 128 B, 16-byte chunks XOR-swizzled by row so that both directions are bank-
// conflict free) and are written back with 8 lanes x 16 B covering one contiguous 128-byte row segment, four rows per
// instruction.  The direct form (each thread stores 16 B of its own row: 32 different lines per warp instruction)
// keeps the LSU busy for ~7-8 k cycles per 128x256 tile — with K = 1536 that is most of the 11-14 k-cycle epilogue, which
// then back-pressures the MMA issuer through the two accumulator stages (tools/gemm_trace.py: 1.5-2.7 k of every
// ~16 k-cycle tile waiting for a free accumulator).
template <int BN, int ACT>
__device__ __forceinline__ void drain_bf16_staged(const EpiParams& ep, uint32_t t_addr, int n_base, float* tbuf, int lane,
                                                  int row, long long orow, int M) {
  constexpr int NC = BN / 64;
  uint8_t* buf = reinterpret_cast<uint8_t*>(tbuf);
  const StagedRows sr = staged_rows(lane, row, orow, M);
  const bool has_bias = ep.bias != nullptr;
  uint32_t r[2][32];
  ptx::tmem_ld32(t_addr, r[0]);
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    ptx::tmem_ld_wait();
    if (c + 1 < NC) ptx::tmem_ld32(t_addr + (c + 1) * 32, r[(c + 1) & 1]);
    const int n0 = n_base + c * 32;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      float v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = __uint_as_float(r[c & 1][j + q]);
      if (has_bias) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(ep.bias + n0 + j));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(ep.bias + n0 + j + 4));
        v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
        v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
      }
      __nv_bfloat162 t0 = __floats2bfloat162_rn(act_fast<ACT>(v[0]), act_fast<ACT>(v[1]));
      __nv_bfloat162 t1 = __floats2bfloat162_rn(act_fast<ACT>(v[2]), act_fast<ACT>(v[3]));
      __nv_bfloat162 t2 = __floats2bfloat162_rn(act_fast<ACT>(v[4]), act_fast<ACT>(v[5]));
      __nv_bfloat162 t3 = __floats2bfloat162_rn(act_fast<ACT>(v[6]), act_fast<ACT>(v[7]));
      uint4 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&t0);
      pk.y = *reinterpret_cast<uint32_t*>(&t1);
      pk.z = *reinterpret_cast<uint32_t*>(&t2);
      pk.w = *reinterpret_cast<uint32_t*>(&t3);
      staged_put(buf, lane, (c & 1) * 4 + (j >> 3), pk);
    }
    // 64 columns of the 32 rows are staged: write them back as whole 128-byte row segments
    if (c & 1) staged_writeback(ep, buf, lane, n0 - 32, sr);
  }
}

// Chunk loop of one accumulator stage with the TMEM load of chunk c+1 in flight while chunk c is processed; kind and
// activation are compile-time so the loop body is straight-line code.
template <int BN, int KIND, int ACT>
__device__ __forceinline__ void drain_loop(const EpiParams& ep, uint32_t t_addr, int n_base, float* tbuf, int lane, int row,
                                           long long orow, int sample, int M, int N) {
  constexpr int NC = BN / 64;
  if (KIND == K_STORE_BF16 && ep.stage_bf16 && n_base + BN / 2 <= N && (ep.ldo & 7) == 0) {
    drain_bf16_staged<BN, ACT>(ep, t_addr, n_base, tbuf, lane, row, orow, M);
    return;
  }
  uint32_t r[2][32];
  ptx::tmem_ld32(t_addr, r[0]);
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    ptx::tmem_ld_wait();
    if (c + 1 < NC) ptx::tmem_ld32(t_addr + (c + 1) * 32, r[(c + 1) & 1]);
    const int n0 = n_base + c * 32;
    if ((KIND == K_GATED || KIND == K_RESID) && ep.prefetch && c + 1 < NC && n0 + 32 < N && row < M)  // next block's lines -> L2
      ptx::prefetch_l2((KIND == K_GATED ? reinterpret_cast<const float*>(ep.out) + orow * ep.ldo : ep.resid + orow * ep.ldr) + n0 + 32);
    if (n0 < N) epilogue_block<KIND, ACT>(ep, r[c & 1], tbuf, lane, row, orow, sample, n0, M, N);
  }
}

// Row bookkeeping of one epilogue thread for tile (m_blk, .): global row (or M when the tile row is padding), remapped
// output row and sample index.
__device__ __forceinline__ void epilogue_row(const EpiParams& ep, const ConvLoad& cv, int m_blk, int quarter, int lane, int M,
                                             int& row, long long& orow, int& sample) {
  const int t = quarter * 32 + lane;
  row = m_blk * BM + t;
  if (cv.enabled) {
    // tile row t -> (frame, x) within the tile; invalid rows (t >= Wt*Nt, x >= W, frame >= n) are dropped
    const int per_group = cv.tiles_x * cv.H;
    const int fg = m_blk / per_group;
    const int rem = m_blk - fg * per_group;
    const int y = rem / cv.tiles_x;
    const int x0 = (rem - y * cv.tiles_x) * cv.Wt;
    const int fi = t / cv.Wt;
    const int x = x0 + (t - fi * cv.Wt);
    const int f = fg * cv.Nt + fi;
    row = (fi < cv.Nt && x < cv.W && f < cv.n_frames) ? (f * cv.H + y) * cv.W + x : M;
  }
  if (row > M) row = M;
  sample = 0;
  orow = epi_out_row(ep, row, sample);
}

// Read-modify-write epilogues (gated fp32 residual, residual + store) walk a warp's 4 column blocks with one DRAM round
// trip per block, and with K = 1536 that chain sets the pace of the out-projections (48-54 % tensor-active,
// profiles/r02_kernels.md).  Each block's residual lines (one 128-byte line per lane = row) are therefore pulled into
// L2 one block ahead: the first block of a tile before the wait for its accumulator (here), block c+1 while block c is
// processed (drain_loop).  (Prefetching the whole NEXT tile instead was measured: +32 % DRAM reads — the lines are
// evicted again before their use — and the K=1536 out-projection 9 % slower.)
template <int BN>
__device__ __forceinline__ void epilogue_prefetch(const EpiParams& ep, int kind, int n_blk, int half, int row, long long orow,
                                                  int M, int N) {
  if ((kind != K_GATED && kind != K_RESID) || !ep.prefetch || row >= M) return;
  const float* base = kind == K_GATED ? reinterpret_cast<const float*>(ep.out) + orow * ep.ldo : ep.resid + orow * ep.ldr;
  const int c0 = n_blk * BN + half * (BN / 2);
  if (c0 < N) ptx::prefetch_l2(base + c0);
}

// Drain one accumulator stage: this warp's 32 rows x (BN/2) columns, TMEM -> registers -> fused epilogue -> global.
template <int BN>
__device__ __forceinline__ void epilogue_drain(const EpiParams& ep, int kind, int n_blk, int quarter, int half, int lane,
                                               float* tbuf, float* tbuf_partner, uint32_t tmem_base, int acc, int row,
                                               long long orow, int sample, int M, int N) {
  const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                          static_cast<uint32_t>(acc * BN + half * (BN / 2));
  if (kind == K_NORM) {
    // the tile holds whole rows (one N tile): pass 1 = sum of squares of this thread's half row, completed with the
    // partner warp (other column half, same TMEM lanes) through the spare column of the transpose tiles; pass 2 =
    // normalise + residual + stores.  TMEM is read twice; the conv output never goes to HBM.
    constexpr int NC = BN / 64;
    uint32_t r[2][32];
    float ss4[4] = {0.f, 0.f, 0.f, 0.f};
    ptx::tmem_ld32(t_addr, r[0]);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      ptx::tmem_ld_wait();
      if (c + 1 < NC) ptx::tmem_ld32(t_addr + (c + 1) * 32, r[(c + 1) & 1]);
#pragma unroll
      for (int j = 0; j < 32; ++j) {  // columns >= N are exact zeros (W rows out of bounds are zero-filled by TMA)
        const float v = __uint_as_float(r[c & 1][j]);
        ss4[j & 3] = fmaf(v, v, ss4[j & 3]);
      }
    }
    float ss = (ss4[0] + ss4[1]) + (ss4[2] + ss4[3]);
    tbuf[lane * 33 + 32] = ss;
    asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
    ss += tbuf_partner[lane * 33 + 32];
    asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");  // both halves have read before a slot is reused
    const float rstd = rsqrtf(ss / static_cast<float>(N) + ep.norm_eps);
    const int n_base = n_blk * BN + half * (BN / 2);
    ptx::tmem_ld32(t_addr, r[0]);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      ptx::tmem_ld_wait();
      if (c + 1 < NC) ptx::tmem_ld32(t_addr + (c + 1) * 32, r[(c + 1) & 1]);
      const int n0 = n_base + c * 32;
      if (n0 < N) epilogue_block_norm(ep, r[c & 1], tbuf, lane, row, orow, rstd, n0, M, N);
    }
    return;
  }
  if (kind == K_QKV) {
    // this warp owns one 128-column head of the tile: pass 1 = sum of squares per row (q/k heads), pass 2 =
    // normalise + rotate + store (staged).  TMEM is read twice; the accumulator never leaves the SM in fp32.  (A
    // one-pass form holding the head row as 64 packed bf16 pairs was measured too: same +0.6 % in-step, 130 B of spills.)
    const int head0 = n_blk * BN + half * (BN / 2);
    const bool is_qk = head0 < ep.qk_cols;
    const float* nw = (head0 < ep.qk_cols / 2) ? ep.qk_wq : ep.qk_wk;
    float rstd = 1.f;
    uint32_t r[2][32];
    if (is_qk && head0 < N) {
      // pass 1, software-pipelined: the TMEM load of chunk c+1 is in flight while chunk c is squared
      float ss[4] = {0.f, 0.f, 0.f, 0.f};
      ptx::tmem_ld32(t_addr, r[0]);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        ptx::tmem_ld_wait();
        if (c + 1 < 4) ptx::tmem_ld32(t_addr + (c + 1) * 32, r[(c + 1) & 1]);
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ep.bias != nullptr) b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + head0 + c * 32 + j));
          const float t0 = __uint_as_float(r[c & 1][j]) + b4.x, t1 = __uint_as_float(r[c & 1][j + 1]) + b4.y;
          const float t2 = __uint_as_float(r[c & 1][j + 2]) + b4.z, t3 = __uint_as_float(r[c & 1][j + 3]) + b4.w;
          ss[0] = fmaf(t0, t0, ss[0]); ss[1] = fmaf(t1, t1, ss[1]); ss[2] = fmaf(t2, t2, ss[2]); ss[3] = fmaf(t3, t3, ss[3]);
        }
      }
      rstd = rsqrtf(((ss[0] + ss[1]) + (ss[2] + ss[3])) * (1.0f / 128.0f) + ep.qk_eps);
    }
    const int tok = row % ep.rows_per_sample;
    const bool row_ok = row < M;
    if (head0 < N) {
      const bool staged = ep.stage_bf16 && (ep.ldo & 7) == 0;
      uint8_t* buf = staged ? reinterpret_cast<uint8_t*>(tbuf) : nullptr;
      StagedRows sr;
      if (staged) sr = staged_rows(lane, row, orow, M);
      ptx::tmem_ld32(t_addr, r[0]);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        ptx::tmem_ld_wait();
        if (c + 1 < 4) ptx::tmem_ld32(t_addr + (c + 1) * 32, r[(c + 1) & 1]);
        epilogue_qkv_chunk(ep, r[c & 1], row_ok, orow, tok, rstd, head0 + c * 32, c * 32, is_qk, nw, buf, lane, c & 1);
        if (staged && (c & 1)) staged_writeback(ep, buf, lane, head0 + (c - 1) * 32, sr);
      }
    }
  } else {
    const int n_base = n_blk * BN + half * (BN / 2);
#define LC_DRAIN(KIND, ACT) drain_loop<BN, KIND, ACT>(ep, t_addr, n_base, tbuf, lane, row, orow, sample, M, N)
#define LC_DRAIN_ACT(KIND)                                   \
  switch (ep.act) {                                          \
    case ACT_GELU_TANH: LC_DRAIN(KIND, ACT_GELU_TANH); break; \
    case ACT_SILU: LC_DRAIN(KIND, ACT_SILU); break;           \
    case ACT_RELU: LC_DRAIN(KIND, ACT_RELU); break;           \
    default: LC_DRAIN(KIND, ACT_NONE); break;                 \
  }
    switch (kind) {
      case K_STORE_BF16: LC_DRAIN_ACT(K_STORE_BF16) break;
      case K_STORE_F32: LC_DRAIN_ACT(K_STORE_F32) break;
      case K_GATED: LC_DRAIN(K_GATED, ACT_NONE); break;
      case K_RESID: LC_DRAIN_ACT(K_RESID) break;
      default: LC_DRAIN_ACT(K_UNPATCH) break;
    }
#undef LC_DRAIN_ACT
#undef LC_DRAIN
  }
}

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmW, int M, int N, int K, int K0, EpiParams ep, TileSched sched,
               ConvLoad cv) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + C::STAGES * C::STAGE_A;
  float* epi_smem = reinterpret_cast<float*>(smem + C::STAGES * C::STAGE);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE + EPI_SMEM);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C::STAGES;
  uint64_t* tmem_full = bars + 2 * C::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = sched.num_m * sched.num_n;
  const int num_kb = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA0);
    ptx::prefetch_tmap(&tmA1);
    ptx::prefetch_tmap(&tmW);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::STAGES; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tmem_full[i], 1);
      ptx::mbar_init(&tmem_empty[i], NUM_EPI_WARPS);  // one arrive per epilogue warp
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc<C::TMEM_COLS>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_grid_sync();  // everything above overlaps the previous kernel's tail; operands are only touched below

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int m_blk, n_blk;
      sched.decode(tile, m_blk, n_blk);
      // implicit-conv tile origin (unused in plain mode)
      int cn0 = 0, cy = 0, cx0 = 0;
      if (cv.enabled) {
        const int tiles_per_row = cv.tiles_x;                 // x tiles per image row
        const int per_group = tiles_per_row * cv.H;            // tiles per frame-group
        const int fg = m_blk / per_group;
        const int rem = m_blk - fg * per_group;
        cy = rem / tiles_per_row;
        cx0 = (rem - cy * tiles_per_row) * cv.Wt;
        cn0 = fg * cv.Nt;
      }
      for (int kb = 0; kb < num_kb; ++kb) {
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        ptx::mbar_expect_tx(&full_bar[stage], cv.enabled ? (cv.a_bytes + C::STAGE_B) : C::STAGE);
        void* a_dst = smem_a + stage * C::STAGE_A;
        void* b_dst = smem_b + stage * C::STAGE_B;
        const int k = kb * BK;
        if (!cv.enabled) {
          if (k < K0)
            ptx::tma_load_2d(a_dst, &tmA0, &full_bar[stage], k, m_blk * BM);
          else
            ptx::tma_load_2d(a_dst, &tmA1, &full_bar[stage], k - K0, m_blk * BM);
          ptx::tma_load_2d(b_dst, &tmW, &full_bar[stage], k, n_blk * BN);
        } else {
          // k-block -> (tap, channel chunk); pole rows read the mirrored tap column for the pad kernel row.
          const int tap = kb / cv.kb_per_tap;
          const int c0 = (kb - tap * cv.kb_per_tap) * BK;
          const int ky = tap / 3;
          int kx = tap - ky * 3;
          if ((cy == 0 && ky == 0) || (cy == cv.H - 1 && ky == 2)) kx = 2 - kx;
          ptx::tma_load_4d(a_dst, &tmA0, &full_bar[stage], c0, cx0 + kx, cy + ky, cn0);
          ptx::tma_load_2d(b_dst, &tmW, &full_bar[stage], k, n_blk * BN);
        }
        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // the whole warp walks the loop converged and one elected lane issues: inside an elect.sync region the compiler
    // keeps the descriptors in uniform registers (1-3 SASS instructions between MMAs instead of a ~20-instruction
    // ELECT / R2UR.BROADCAST sequence per MMA)
    constexpr uint32_t idesc = ptx::make_idesc_bf16(BM, BN, 0, 0);
    constexpr uint32_t desc_hi = ptx::smem_desc_hi(1024);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        const uint32_t a_lo = ptx::smem_desc_lo(ptx::smem_u32(smem_a + stage * C::STAGE_A), 16);
        const uint32_t b_lo = ptx::smem_desc_lo(ptx::smem_u32(smem_b + stage * C::STAGE_B), 16);
        if (ptx::elect_one()) {
#pragma unroll
          for (int kk = 0; kk < BK / UMMA_K; ++kk)
            ptx::umma_f16_lh(d_tmem, a_lo + ((kk * UMMA_K * 2) >> 4), b_lo + ((kk * UMMA_K * 2) >> 4), desc_hi, idesc,
                             (kb | kk) != 0 ? 1u : 0u);
          ptx::umma_commit(&empty_bar[stage]);
          if (kb == num_kb - 1) ptx::umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= EPI_WARP0) {
    // ===================== epilogue (8 warps) =====================
    const int ew = warp - EPI_WARP0;
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access (= warp id % 4)
    const int half = ew >> 2;      // which half of the tile's columns
    float* tbuf = epi_smem + ew * (32 * 33);
    float* tbuf_partner = epi_smem + (ew ^ 4) * (32 * 33);
    const int kind = epi_kind(ep);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int m_blk, n_blk;
      sched.decode(tile, m_blk, n_blk);
      int row, sample;
      long long orow;
      epilogue_row(ep, cv, m_blk, quarter, lane, M, row, orow, sample);
      epilogue_prefetch<BN>(ep, kind, n_blk, half, row, orow, M, N);
      ptx::mbar_wait(&tmem_full[acc], acc_phase);
      ptx::tc_fence_after();
      epilogue_drain<BN>(ep, kind, n_blk, quarter, half, lane, tbuf, tbuf_partner, tmem_base, acc, row, orow, sample, M, N);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------------ CTA-pair variant
// cta_group::2: a cluster of two CTAs computes a 256 x 256 tile.  Each CTA loads ITS 128 rows of A and HALF of the W
// tile (128 of the 256 N rows) — 32 KB per k-block instead of 48 KB, which takes the shared-memory port below the
// rate the tensor pipe consumes operands at — the leader CTA issues one M = 256 tcgen05.mma per K-step, and each
// CTA's TMEM receives the accumulators of its own 128 rows.  Barriers: TMA bytes of both CTAs are signalled on the
// leader's `full` barrier; tcgen05.commit multicasts to the `empty` / `tmem_full` barriers of both CTAs; the
// epilogue warps of both CTAs arrive on the leader's `tmem_empty`.
struct Cfg2 {
  static constexpr int BN = 256;
  static constexpr int STAGE_A = BM * BK * 2;        // 16 KB: this CTA's 128 rows of A
  static constexpr int STAGE_B = (BN / 2) * BK * 2;  // 16 KB: this CTA's half of the W tile
  static constexpr int STAGE = STAGE_A + STAGE_B;
  static constexpr int STAGES = 6;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int SMEM = STAGES * STAGE + EPI_SMEM + 1024 + 256;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                const __grid_constant__ CUtensorMap tmW, int M, int N, int K, int K0, EpiParams ep, TileSched sched,
                ConvLoad cv) {
  using C = Cfg2;
  constexpr int BN = C::BN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + C::STAGES * C::STAGE_A;
  float* epi_smem = reinterpret_cast<float*>(smem + C::STAGES * C::STAGE);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE + EPI_SMEM);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C::STAGES;
  uint64_t* tmem_full = bars + 2 * C::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = ptx::cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int num_pairs = static_cast<int>(gridDim.x >> 1);
  const int pair_id = static_cast<int>(blockIdx.x >> 1);
  const int num_tiles = sched.num_m * sched.num_n;  // sched.num_m counts 256-row pair tiles
  const int num_kb = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA0);
    ptx::prefetch_tmap(&tmA1);
    ptx::prefetch_tmap(&tmW);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::STAGES; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tmem_full[i], 1);
      ptx::mbar_init(&tmem_empty[i], 2 * NUM_EPI_WARPS);  // epilogue warps of BOTH CTAs
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc_pair<C::TMEM_COLS>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();  // peer's barriers are initialised and its TMEM is allocated
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_grid_sync();  // everything above overlaps the previous kernel's tail; operands are only touched below

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer (both CTAs) =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = pair_id; tile < num_tiles; tile += num_pairs) {
      int mp, n_blk;
      sched.decode(tile, mp, n_blk);
      const int m_blk = mp * 2 + static_cast<int>(cta_rank);
      int cn0 = 0, cy = 0, cx0 = 0;
      if (cv.enabled) {
        const int per_group = cv.tiles_x * cv.H;
        const int fg = m_blk / per_group;
        const int rem = m_blk - fg * per_group;
        cy = rem / cv.tiles_x;
        cx0 = (rem - cy * cv.tiles_x) * cv.Wt;
        cn0 = fg * cv.Nt;
      }
      for (int kb = 0; kb < num_kb; ++kb) {
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        const uint32_t full_leader = ptx::map_to_cta(&full_bar[stage], 0);
        if (leader) {
          // bytes of both CTAs; in conv mode the partner's A box may be a different (but equally sized) tile
          ptx::mbar_expect_tx(&full_bar[stage], 2u * (cv.enabled ? (cv.a_bytes + C::STAGE_B) : C::STAGE));
        }
        void* a_dst = smem_a + stage * C::STAGE_A;
        void* b_dst = smem_b + stage * C::STAGE_B;
        const int k = kb * BK;
        if (!cv.enabled) {
          if (k < K0)
            ptx::tma_load_2d_pair(a_dst, &tmA0, full_leader, k, m_blk * BM);
          else
            ptx::tma_load_2d_pair(a_dst, &tmA1, full_leader, k - K0, m_blk * BM);
        } else {
          const int tap = kb / cv.kb_per_tap;
          const int c0 = (kb - tap * cv.kb_per_tap) * BK;
          const int ky = tap / 3;
          int kx = tap - ky * 3;
          if ((cy == 0 && ky == 0) || (cy == cv.H - 1 && ky == 2)) kx = 2 - kx;
          ptx::tma_load_4d_pair(a_dst, &tmA0, full_leader, c0, cx0 + kx, cy + ky, cn0);
        }
        ptx::tma_load_2d_pair(b_dst, &tmW, full_leader, k, n_blk * BN + static_cast<int>(cta_rank) * (BN / 2));
        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && leader) {
    // ===================== MMA issuer (leader CTA only) =====================
    // converged warp + elect.sync-guarded issue (uniform-datapath descriptors), as in the single-CTA kernel
    constexpr uint32_t idesc = ptx::make_idesc_bf16(2 * BM, BN, 0, 0);
    constexpr uint32_t desc_hi = ptx::smem_desc_hi(1024);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    long long* trace = (pair_id == 0 && lane == 0) ? g_gemm_trace : nullptr;
    int tix = 0;
    for (int tile = pair_id; tile < num_tiles; tile += num_pairs, ++tix) {
      if (trace != nullptr && tix < 64) trace[tix * 4 + 0] = clock64();
      ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      ptx::tc_fence_after();
      if (trace != nullptr && tix < 64) trace[tix * 4 + 1] = clock64();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        if (trace != nullptr && tix < 64 && kb == 0) trace[tix * 4 + 2] = clock64();
        const uint32_t a_lo = ptx::smem_desc_lo(ptx::smem_u32(smem_a + stage * C::STAGE_A), 16);
        const uint32_t b_lo = ptx::smem_desc_lo(ptx::smem_u32(smem_b + stage * C::STAGE_B), 16);
        if (ptx::elect_one()) {
#pragma unroll
          for (int kk = 0; kk < BK / UMMA_K; ++kk)
            ptx::umma_f16_pair_lh(d_tmem, a_lo + ((kk * UMMA_K * 2) >> 4), b_lo + ((kk * UMMA_K * 2) >> 4), desc_hi, idesc,
                                  (kb | kk) != 0 ? 1u : 0u);
          ptx::umma_commit_pair(&empty_bar[stage]);
          if (kb == num_kb - 1) ptx::umma_commit_pair(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
      }
      if (trace != nullptr && tix < 64) trace[tix * 4 + 3] = clock64();
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= EPI_WARP0) {
    // ===================== epilogue (8 warps per CTA, own 128 rows) =====================
    const int ew = warp - EPI_WARP0;
    const int quarter = warp & 3;
    const int half = ew >> 2;
    float* tbuf = epi_smem + ew * (32 * 33);
    float* tbuf_partner = epi_smem + (ew ^ 4) * (32 * 33);
    const int kind = epi_kind(ep);
    int acc = 0;
    uint32_t acc_phase = 0;
    long long* trace = (pair_id == 0 && leader && ew == 0 && lane == 0) ? g_gemm_trace : nullptr;
    int tix = 0;
    for (int tile = pair_id; tile < num_tiles; tile += num_pairs, ++tix) {
      int mp, n_blk;
      sched.decode(tile, mp, n_blk);
      const int m_blk = mp * 2 + static_cast<int>(cta_rank);
      int row, sample;
      long long orow;
      epilogue_row(ep, cv, m_blk, quarter, lane, M, row, orow, sample);
      epilogue_prefetch<BN>(ep, kind, n_blk, half, row, orow, M, N);
      if (trace != nullptr && tix < 64) trace[256 + tix * 4 + 0] = clock64();
      ptx::mbar_wait(&tmem_full[acc], acc_phase);
      ptx::tc_fence_after();
      if (trace != nullptr && tix < 64) trace[256 + tix * 4 + 1] = clock64();

      epilogue_drain<BN>(ep, kind, n_blk, quarter, half, lane, tbuf, tbuf_partner, tmem_base, acc, row, orow, sample, M, N);
      ptx::tc_fence_before();
      __syncwarp();
      if (trace != nullptr && tix < 64) trace[256 + tix * 4 + 2] = clock64();
      if (lane == 0) ptx::mbar_arrive_cluster(ptx::map_to_cta(&tmem_empty[acc], 0));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();  // nobody may exit (or free TMEM) while the partner still reads its smem / signals its barriers
  if (warp == 2) ptx::tmem_dealloc_pair<C::TMEM_COLS>(tmem_base);
}

// algorithmic HBM bytes of one launch: A once (an implicit conv reads its input once, not 9 times), W once, the output
// once (read-modify-write epilogues: twice)
double algorithmic_bytes(int M, int N, int K, const EpiParams& ep, const ConvLoad& cv) {
  const double a = 2.0 * M * (cv.enabled ? K / 9.0 : static_cast<double>(K));
  const double out_b = ep.mode == EPI_NORM_RESID ? (ep.xres ? 8.0 : 0.0) + (ep.out ? 2.0 : 0.0)
                       : (ep.mode == EPI_GATED_RESID || ep.mode == EPI_RESID_STORE) ? 8.0
                       : (ep.mode == EPI_UNPATCHIFY || ep.out_f32) ? 4.0 : 2.0;
  const double n_out = ep.mode == EPI_UNPATCHIFY ? ep.n_valid : N;
  return a + 2.0 * N * K + out_b * M * n_out;
}

int launch_pair(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w, int M_tiles, int M, int N, int K,
                int K0, const EpiParams& ep, const ConvLoad& cv, cudaStream_t stream) {
  using C = Cfg2;
  static PerDevice<bool> attr_set;
  if (!attr_set.here()) {
    LC_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    attr_set.here() = true;
  }
  TileSched sched;
  sched.num_m = ceil_div(M_tiles, 2);  // 256-row pair tiles
  sched.num_n = ceil_div(N, C::BN);
  static const int group_m_env = [] { const char* e = getenv("LADCAST_B200_GROUP_M"); return e ? atoi(e) : 0; }();
  sched.group_m = group_m_env > 0 ? group_m_env : 4;
  const int tiles = sched.num_m * sched.num_n;
  const int pairs = num_sms() / 2;
  const int grid = 2 * (tiles < pairs ? tiles : pairs);
  const int cls = cv.enabled ? PROF_CONV : PROF_GEMM;
  const double flops = 2.0 * M * N * (cv.enabled ? 9.0 * cv.c_real : static_cast<double>(K));
  // block-ahead L2 prefetch of residual lines: measured neutral in-step (17.5-18.0 ms per 375M call either way), so off
  static const bool epi_prefetch = [] { const char* e = getenv("LADCAST_B200_EPI_PREFETCH"); return e != nullptr && e[0] == '1'; }();
  static const bool epi_stage = [] { const char* e = getenv("LADCAST_B200_EPI_STAGE"); return !(e != nullptr && e[0] == '0'); }();
  EpiParams epl = ep;
  epl.prefetch = epi_prefetch ? 1 : 0;
  epl.stage_bf16 = epi_stage ? 1 : 0;
  prof_begin(cls, stream);
  LC_CHECK_CUDA(launch_kernel(gemm_tc2_kernel, grid, NUM_THREADS, C::SMEM, stream, a0, a1, w, M, N, K, K0, epl, sched, cv));
  prof_end(cls, flops, stream, algorithmic_bytes(M, N, K, ep, cv));
  LC_LAUNCH_CHECK();
  return 0;
}

template <int BN>
int launch(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w, int M_tiles, int M, int N, int K, int K0,
           const EpiParams& ep, const ConvLoad& cv, cudaStream_t stream) {
  using C = Cfg<BN>;
  static PerDevice<bool> attr_set;
  if (!attr_set.here()) {
    LC_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    attr_set.here() = true;
  }
  TileSched sched;
  sched.num_m = M_tiles;
  sched.num_n = ceil_div(N, BN);
  sched.group_m = 8;
  const int tiles = sched.num_m * sched.num_n;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  const int cls = cv.enabled ? PROF_CONV : PROF_GEMM;
  const double flops = 2.0 * M * N * (cv.enabled ? 9.0 * cv.c_real : static_cast<double>(K));
  prof_begin(cls, stream);
  LC_CHECK_CUDA(launch_kernel(gemm_tc_kernel<BN>, grid, NUM_THREADS, C::SMEM, stream, a0, a1, w, M, N, K, K0, ep, sched, cv));
  prof_end(cls, flops, stream, algorithmic_bytes(M, N, K, ep, cv));
  LC_LAUNCH_CHECK();
  return 0;
}

}  // namespace

int num_sms() {
  static PerDevice<int> cache;
  int& n = cache.here();
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

static int pick_bn(int N) { return (N > 128) ? 256 : 128; }

static int check_norm_epilogue(const EpiParams& ep, int N, int bn) {
  if (ep.mode != EPI_NORM_RESID) return 0;
  LC_REQUIRE(N <= bn && N % 4 == 0, "fused RMSNorm epilogue needs all output channels in one N tile (N <= 256, N % 4 == 0)");
  LC_REQUIRE(ep.bias == nullptr && ep.norm_w != nullptr && ep.norm_b != nullptr, "fused RMSNorm epilogue: no GEMM bias, norm weight + bias required");
  LC_REQUIRE(ep.xres == nullptr || ep.ldr % 4 == 0, "fused RMSNorm epilogue: residual row pitch must be a multiple of 4");
  LC_REQUIRE(ep.out == nullptr || ep.ldo % 4 == 0, "fused RMSNorm epilogue: output row pitch must be a multiple of 4");
  return 0;
}

// CTA pairs pay off once there are enough 256 x 256 tiles to fill the 74 pairs; LADCAST_B200_GEMM_PAIR=0 disables.
static bool use_pair(int m_tiles, int N) {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("LADCAST_B200_GEMM_PAIR");
    mode = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return mode == 1 && m_tiles >= 2 && static_cast<long long>(ceil_div(m_tiles, 2)) * ceil_div(N, 256) >= 37;
}

int gemm_set_trace(long long* buf) {
  LC_CHECK_CUDA(cudaMemcpyToSymbol(g_gemm_trace, &buf, sizeof(buf)));
  return 0;
}

int gemm_bf16(const GemmArgs& g, cudaStream_t stream) {
  LC_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0, "empty GEMM");
  LC_REQUIRE(g.epi.qk_cols == 0 || (g.N > 128 && g.N % 128 == 0 && g.epi.qk_cols % 256 == 0 && !g.epi.out_f32),
             "fused qk-norm/rope epilogue needs head-aligned (128) columns and bf16 output");
  const int K0 = (g.A1 != nullptr) ? g.K0 : g.K;
  LC_REQUIRE(g.A1 == nullptr || (K0 % BK == 0 && K0 > 0 && K0 < g.K), "split-K source boundary must be a multiple of 64");
  const int bn = pick_bn(g.N);
  LC_TRY(check_norm_epilogue(g.epi, g.N, bn));
  CUtensorMap ta0, ta1, tw;
  LC_TRY(make_tmap_2d_bf16(&ta0, g.A0, static_cast<uint64_t>(K0), static_cast<uint64_t>(g.M),
                           static_cast<uint64_t>(g.lda0) * 2, BK, BM));
  if (g.A1 != nullptr)
    LC_TRY(make_tmap_2d_bf16(&ta1, g.A1, static_cast<uint64_t>(g.K - K0), static_cast<uint64_t>(g.M),
                             static_cast<uint64_t>(g.lda1) * 2, BK, BM));
  else
    ta1 = ta0;
  LC_TRY(make_tmap_2d_bf16(&tw, g.W, static_cast<uint64_t>(g.K), static_cast<uint64_t>(g.N),
                           static_cast<uint64_t>(g.ldw) * 2, BK, static_cast<uint32_t>(bn)));
  ConvLoad cv;
  const int m_tiles = ceil_div(g.M, BM);
  if (bn == 256 && use_pair(m_tiles, g.N)) {
    // pair kernel: each CTA loads half of the W tile -> its tensor map has a 128-row box
    LC_TRY(make_tmap_2d_bf16(&tw, g.W, static_cast<uint64_t>(g.K), static_cast<uint64_t>(g.N),
                             static_cast<uint64_t>(g.ldw) * 2, BK, 128));
    return launch_pair(ta0, ta1, tw, m_tiles, g.M, g.N, g.K, K0, g.epi, cv, stream);
  }
  if (bn == 256) return launch<256>(ta0, ta1, tw, m_tiles, g.M, g.N, g.K, K0, g.epi, cv, stream);
  return launch<128>(ta0, ta1, tw, m_tiles, g.M, g.N, g.K, K0, g.epi, cv, stream);
}

// Implicit-GEMM 3x3 sphere convolution.  xpad: [n, H+2, W+2, Cp] bf16 (sphere-padded, channel-padded to a
// multiple of 64), wmat: [C_out, 9*Cp] bf16 (tap-major: k = (ky*3+kx)*Cp + c).  Output rows are pixels
// (f*H + y)*W + x, columns are output channels; epilogue as for GEMMs.
int conv3x3_bf16(const void* xpad, int n_frames, int H, int W, int Cp, const void* wmat, int C_out,
                 const EpiParams& epi, cudaStream_t stream, int c_real) {
  LC_REQUIRE(Cp % BK == 0, "conv input channels must be padded to a multiple of 64");
  ConvLoad cv;
  cv.enabled = 1;
  cv.H = H;
  cv.W = W;
  cv.n_frames = n_frames;
  // tile = Nt frames x 1 row x Wt columns, Wt*Nt <= 128
  if (W >= 128) { cv.Wt = (W % 120 == 0) ? 120 : 128; cv.Nt = 1; }
  else { cv.Wt = W; cv.Nt = 128 / W; if (cv.Nt > n_frames) cv.Nt = n_frames; if (cv.Nt < 1) cv.Nt = 1; }
  cv.tiles_x = ceil_div(W, cv.Wt);
  cv.kb_per_tap = Cp / BK;
  cv.a_bytes = BK * 2 * cv.Wt * cv.Nt;
  cv.c_real = c_real > 0 ? c_real : Cp;
  const int frame_groups = ceil_div(n_frames, cv.Nt);
  const int m_tiles = frame_groups * H * cv.tiles_x;
  const int M = n_frames * H * W;
  const int K = 9 * Cp;
  const int bn = pick_bn(C_out);
  LC_TRY(check_norm_epilogue(epi, C_out, bn));
  CUtensorMap ta, tw;
  uint64_t dims[4] = {static_cast<uint64_t>(Cp), static_cast<uint64_t>(W + 2), static_cast<uint64_t>(H + 2),
                      static_cast<uint64_t>(n_frames)};
  uint64_t strides[3] = {static_cast<uint64_t>(Cp) * 2, static_cast<uint64_t>(Cp) * 2 * (W + 2),
                         static_cast<uint64_t>(Cp) * 2 * (W + 2) * (H + 2)};
  uint32_t box[4] = {BK, static_cast<uint32_t>(cv.Wt), 1, static_cast<uint32_t>(cv.Nt)};
  LC_TRY(make_tmap_bf16(&ta, xpad, 4, dims, strides, box));
  LC_TRY(make_tmap_2d_bf16(&tw, wmat, static_cast<uint64_t>(K), static_cast<uint64_t>(C_out),
                           static_cast<uint64_t>(K) * 2, BK, static_cast<uint32_t>(bn)));
  if (bn == 256 && use_pair(m_tiles, C_out)) {
    LC_TRY(make_tmap_2d_bf16(&tw, wmat, static_cast<uint64_t>(K), static_cast<uint64_t>(C_out),
                             static_cast<uint64_t>(K) * 2, BK, 128));
    return launch_pair(ta, ta, tw, m_tiles, M, C_out, K, K, epi, cv, stream);
  }
  if (bn == 256) return launch<256>(ta, ta, tw, m_tiles, M, C_out, K, K, epi, cv, stream);
  return launch<128>(ta, ta, tw, m_tiles, M, C_out, K, K, epi, cv, stream);
}

}  // namespace lc
