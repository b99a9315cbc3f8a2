// FP32 SIMT GEMM — the "FP32 validation mode" of the denoiser/decoder (north star: rel-L2 <= 1e-4 vs the oracle).
// Same problem statement and epilogues as the tensor-core GEMM (common.cuh), plain shared-memory tiling.
#include "common.cuh"
#include "dcae_kernels.h"

namespace lc {
namespace {

constexpr int TM = 64, TN = 64, TK = 16;

__device__ __forceinline__ void epi_one(const EpiParams& ep, float v, int row, int n, int N) {
  int sample;
  const long long orow = epi_out_row(ep, row, sample);
  if (ep.bias) v += __ldg(ep.bias + n);
  v = apply_act(v, ep.act);
  switch (ep.mode) {
    case EPI_STORE:
      if (ep.out_f32) reinterpret_cast<float*>(ep.out)[orow * ep.ldo + n] = v;
      else reinterpret_cast<bf16*>(ep.out)[orow * ep.ldo + n] = __float2bfloat16_rn(v);
      break;
    case EPI_RESID_STORE:
      v += ep.resid[orow * ep.ldr + n];
      if (ep.out_f32) reinterpret_cast<float*>(ep.out)[orow * ep.ldo + n] = v;
      else reinterpret_cast<bf16*>(ep.out)[orow * ep.ldo + n] = __float2bfloat16_rn(v);
      break;
    case EPI_GATED_RESID: {
      const float g = ep.gate ? __ldg(ep.gate + static_cast<long long>(sample) * ep.gate_stride + n) : 1.f;
      reinterpret_cast<float*>(ep.out)[orow * ep.ldo + n] += g * v;
      break;
    }
    case EPI_UNPATCHIFY:
      if (n < ep.n_valid) {
        const int r_in = row - sample * ep.rows_per_sample;
        if (ep.ch_scale != nullptr) v = v * __ldg(ep.ch_scale + n) + __ldg(ep.ch_shift + n);
        long long chs;
        const long long base = epi_unpatch_base(ep, sample, chs);
        reinterpret_cast<float*>(ep.out)[base + n * chs + r_in] = v;
      }
      break;
  }
}

struct ConvF32 {
  int enabled = 0, H = 0, W = 0, Cp = 0;
};

__global__ void __launch_bounds__(256) gemm_f32_kernel(const float* __restrict__ A0, long long lda0, int K0,
                                                       const float* __restrict__ A1, long long lda1,
                                                       const float* __restrict__ W, long long ldw, int M, int N, int K,
                                                       EpiParams ep, ConvF32 cv) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Ws[TK][TN + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += TK) {
    // 64x16 tiles, 256 threads -> 4 elements each
    for (int i = threadIdx.x; i < TM * TK; i += 256) {
      const int r = i / TK, c = i % TK;
      const int gr = m0 + r, gk = k0 + c;
      float v = 0.f;
      if (gr < M && gk < K) {
        if (!cv.enabled) {
          v = (gk < K0) ? A0[gr * lda0 + gk] : A1[gr * lda1 + (gk - K0)];
        } else {
          // implicit 3x3 sphere conv: row -> (frame, y, x); k -> (tap, channel); pole rows mirror the pad kernel row
          const int x = gr % cv.W, y = (gr / cv.W) % cv.H, f = gr / (cv.W * cv.H);
          const int tap = gk / cv.Cp, ch = gk - tap * cv.Cp;
          const int ky = tap / 3;
          int kx = tap - ky * 3;
          if ((y == 0 && ky == 0) || (y == cv.H - 1 && ky == 2)) kx = 2 - kx;
          v = A0[((static_cast<long long>(f) * (cv.H + 2) + y + ky) * (cv.W + 2) + x + kx) * cv.Cp + ch];
        }
      }
      As[c][r] = v;
      const int gn = n0 + r;
      Ws[c][r] = (gn < N && gk < K) ? W[gn * ldw + gk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Ws[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + ty * 4 + i;
    if (row >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < N) epi_one(ep, acc[i][j], row, n, N);
    }
  }
}

}  // namespace

int gemm_f32(const GemmArgs& g, cudaStream_t stream) {
  LC_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0, "empty GEMM");
  LC_REQUIRE(g.epi.mode != EPI_NORM_RESID, "the fused RMSNorm epilogue exists on the tensor-core path only");
  const int K0 = (g.A1 != nullptr) ? g.K0 : g.K;
  dim3 grid(ceil_div(g.N, TN), ceil_div(g.M, TM));
  gemm_f32_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float*>(g.A0), g.lda0, K0,
                                            reinterpret_cast<const float*>(g.A1), g.lda1,
                                            reinterpret_cast<const float*>(g.W), g.ldw, g.M, g.N, g.K, g.epi,
                                            ConvF32());
  LC_LAUNCH_CHECK();
  return 0;
}

int conv3x3_f32(const float* xpad, int n_frames, int H, int W, int Cp, const float* wmat, int C_out,
                const EpiParams& epi, cudaStream_t stream) {
  LC_REQUIRE(epi.mode != EPI_NORM_RESID, "the fused RMSNorm epilogue exists on the tensor-core path only");
  ConvF32 cv;
  cv.enabled = 1; cv.H = H; cv.W = W; cv.Cp = Cp;
  const int M = n_frames * H * W, K = 9 * Cp;
  dim3 grid(ceil_div(C_out, TN), ceil_div(M, TM));
  gemm_f32_kernel<<<grid, 256, 0, stream>>>(xpad, 0, K, nullptr, 0, wmat, K, M, C_out, K, epi, cv);
  LC_LAUNCH_CHECK();
  return 0;
}

}  // namespace lc
