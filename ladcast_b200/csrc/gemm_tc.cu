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
#include "common.cuh"
#include "gemm_tc.h"
#include "ptx.cuh"
#include "tmap.h"

namespace lc {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 256;
constexpr int EPI_WARP0 = 4;

template <int BN>
struct Cfg {
  static constexpr int STAGE_A = BM * BK * 2;
  static constexpr int STAGE_B = BN * BK * 2;
  static constexpr int STAGE = STAGE_A + STAGE_B;
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int TMEM_COLS = 2 * BN;  // two accumulator stages
  static constexpr int SMEM = STAGES * STAGE + 1024 /*align slack*/ + 256 /*barriers*/;
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

template <int BN>
__device__ __forceinline__ void epilogue_chunk(const EpiParams& ep, const uint32_t (&r)[32], int row, int n0, int M,
                                               int N) {
  if (row >= M) return;
  int sample;
  const long long orow = epi_out_row(ep, row, sample);
  const bool full = (n0 + 32 <= N);
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
  if (ep.bias != nullptr) {
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + n0 + j));
        v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n0 + j < N) v[j] += __ldg(ep.bias + n0 + j);
    }
  }
  if (ep.act != ACT_NONE) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], ep.act);
  }
  switch (ep.mode) {
    case EPI_STORE:
    case EPI_RESID_STORE: {
      if (ep.mode == EPI_RESID_STORE) {
        const float* rp = ep.resid + orow * ep.ldr + n0;
        if (full) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 q = *reinterpret_cast<const float4*>(rp + j);
            v[j] += q.x; v[j + 1] += q.y; v[j + 2] += q.z; v[j + 3] += q.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n0 + j < N) v[j] += rp[j];
        }
      }
      if (ep.out_f32) {
        float* op = reinterpret_cast<float*>(ep.out) + orow * ep.ldo + n0;
        if (full) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(op + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n0 + j < N) op[j] = v[j];
        }
      } else {
        bf16* op = reinterpret_cast<bf16*>(ep.out) + orow * ep.ldo + n0;
        if (full) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 pk;
            __nv_bfloat162 t0 = __floats2bfloat162_rn(v[j], v[j + 1]);
            __nv_bfloat162 t1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
            __nv_bfloat162 t2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]);
            __nv_bfloat162 t3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
            pk.x = *reinterpret_cast<uint32_t*>(&t0);
            pk.y = *reinterpret_cast<uint32_t*>(&t1);
            pk.z = *reinterpret_cast<uint32_t*>(&t2);
            pk.w = *reinterpret_cast<uint32_t*>(&t3);
            *reinterpret_cast<uint4*>(op + j) = pk;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n0 + j < N) op[j] = __float2bfloat16_rn(v[j]);
        }
      }
      break;
    }
    case EPI_GATED_RESID: {
      float* op = reinterpret_cast<float*>(ep.out) + orow * ep.ldo + n0;
      const float* gp = ep.gate ? ep.gate + static_cast<long long>(sample) * ep.gate_stride + n0 : nullptr;
      if (full) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 h = *reinterpret_cast<float4*>(op + j);
          float4 g = gp ? __ldg(reinterpret_cast<const float4*>(gp + j)) : make_float4(1.f, 1.f, 1.f, 1.f);
          h.x += g.x * v[j]; h.y += g.y * v[j + 1]; h.z += g.z * v[j + 2]; h.w += g.w * v[j + 3];
          *reinterpret_cast<float4*>(op + j) = h;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (n0 + j < N) op[j] += (gp ? __ldg(gp + j) : 1.f) * v[j];
      }
      break;
    }
    case EPI_UNPATCHIFY: {
      const int r_in = row - sample * ep.rows_per_sample;
      float* op = reinterpret_cast<float*>(ep.out) +
                  (static_cast<long long>(sample) * ep.n_valid + n0) * ep.rows_per_sample + r_in;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n0 + j < ep.n_valid) {
          float o = v[j];
          if (ep.ch_scale != nullptr) o = o * __ldg(ep.ch_scale + n0 + j) + __ldg(ep.ch_shift + n0 + j);
          op[static_cast<long long>(j) * ep.rows_per_sample] = o;
        }
      break;
    }
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
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE);
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
      ptx::mbar_init(&tmem_empty[i], 4);  // one arrive per epilogue warp
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc<C::TMEM_COLS>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

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
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = ptx::make_idesc_bf16(BM, BN, 0, 0);
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
        const uint32_t a_base = ptx::smem_u32(smem_a + stage * C::STAGE_A);
        const uint32_t b_base = ptx::smem_u32(smem_b + stage * C::STAGE_B);
#pragma unroll
        for (int kk = 0; kk < BK / UMMA_K; ++kk) {
          const uint64_t da = ptx::make_smem_desc(a_base + kk * UMMA_K * 2, 16, 1024);
          const uint64_t db = ptx::make_smem_desc(b_base + kk * UMMA_K * 2, 16, 1024);
          ptx::umma_f16(d_tmem, da, db, idesc, (kb | kk) != 0 ? 1u : 0u);
        }
        ptx::umma_commit(&empty_bar[stage]);
        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
      }
      ptx::umma_commit(&tmem_full[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= EPI_WARP0) {
    // ===================== epilogue =====================
    const int quarter = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int m_blk, n_blk;
      sched.decode(tile, m_blk, n_blk);
      ptx::mbar_wait(&tmem_full[acc], acc_phase);
      ptx::tc_fence_after();
      const int t = quarter * 32 + lane;
      int row = m_blk * BM + t;
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
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * BN);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        ptx::tmem_ld32(t_addr + c * 32, r);
        ptx::tmem_ld_wait();
        const int n0 = n_blk * BN + c * 32;
        if (n0 < N) epilogue_chunk<BN>(ep, r, row, n0, M, N);
      }
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

template <int BN>
int launch(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w, int M_tiles, int M, int N, int K, int K0,
           const EpiParams& ep, const ConvLoad& cv, cudaStream_t stream) {
  using C = Cfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    LC_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    attr_set = true;
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
  gemm_tc_kernel<BN><<<grid, NUM_THREADS, C::SMEM, stream>>>(a0, a1, w, M, N, K, K0, ep, sched, cv);
  prof_end(cls, flops, stream);
  LC_LAUNCH_CHECK();
  return 0;
}

}  // namespace

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

static int pick_bn(int N) { return (N > 128) ? 256 : 128; }

int gemm_bf16(const GemmArgs& g, cudaStream_t stream) {
  LC_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0, "empty GEMM");
  const int K0 = (g.A1 != nullptr) ? g.K0 : g.K;
  LC_REQUIRE(g.A1 == nullptr || (K0 % BK == 0 && K0 > 0 && K0 < g.K), "split-K source boundary must be a multiple of 64");
  const int bn = pick_bn(g.N);
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
  CUtensorMap ta, tw;
  uint64_t dims[4] = {static_cast<uint64_t>(Cp), static_cast<uint64_t>(W + 2), static_cast<uint64_t>(H + 2),
                      static_cast<uint64_t>(n_frames)};
  uint64_t strides[3] = {static_cast<uint64_t>(Cp) * 2, static_cast<uint64_t>(Cp) * 2 * (W + 2),
                         static_cast<uint64_t>(Cp) * 2 * (W + 2) * (H + 2)};
  uint32_t box[4] = {BK, static_cast<uint32_t>(cv.Wt), 1, static_cast<uint32_t>(cv.Nt)};
  LC_TRY(make_tmap_bf16(&ta, xpad, 4, dims, strides, box));
  LC_TRY(make_tmap_2d_bf16(&tw, wmat, static_cast<uint64_t>(K), static_cast<uint64_t>(C_out),
                           static_cast<uint64_t>(K) * 2, BK, static_cast<uint32_t>(bn)));
  if (bn == 256) return launch<256>(ta, ta, tw, m_tiles, M, C_out, K, K, epi, cv, stream);
  return launch<128>(ta, ta, tw, m_tiles, M, C_out, K, K, epi, cv, stream);
}

}  // namespace lc
