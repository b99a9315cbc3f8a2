// Ensemble forecast metrics on the device: latitude-weighted ensemble-mean squared error, CRPS skill / spread /
// total.  Reference: evaluate/utils.py:40-118 (pointwise functions) and the assembly loop
// evaluate/evaluate_ens_gpu.py:339-415 (weights are float64, the SST channel is reduced with nanmean).
//
//   fields [M, N, HW] f32  (M members, N = (channel, lead) planes, HW pixels)     truth [N, HW] f32 (NaN allowed)
// One thread per pixel: the M member values are staged in shared memory (column per thread), the spread is the
// mean absolute difference over all member pairs — algebraically identical to the reference's sorted formula
// 2/(M(M-1)) * sum_i (2i - M - 1) x_(i).  Per-pixel values are fp32 (as in the reference), the latitude-weighted
// spatial sums are fp64: warp-shuffle + shared-memory block reduction, one fp64 atomicAdd per block and metric.
#include "../../include/ladcast_b200.h"
#include "common.cuh"

namespace lc {
namespace {

constexpr int THREADS = 256;

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <bool REDUCE>
__global__ void __launch_bounds__(THREADS) metrics_kernel(const float* __restrict__ fields, const float* __restrict__ truth,
                                                          const double* __restrict__ latw, int M, long long N, int H,
                                                          int W, double* __restrict__ sums, double* __restrict__ counts,
                                                          float* __restrict__ out_skill, float* __restrict__ out_spread,
                                                          float* __restrict__ out_mean) {
  extern __shared__ float xs[];  // [M][THREADS]
  const int HW = H * W;
  const long long n = blockIdx.y;
  const int p = blockIdx.x * THREADS + threadIdx.x;
  const bool active = p < HW;
  float msum = 0.f, skill = 0.f, spread = 0.f, y = 0.f;
  if (active) {
    y = truth != nullptr ? truth[n * HW + p] : 0.f;
    for (int m = 0; m < M; ++m) {
      const float x = fields[(static_cast<long long>(m) * N + n) * HW + p];
      xs[m * THREADS + threadIdx.x] = x;
      msum += x;
      skill += fabsf(y - x);
    }
    float pair = 0.f;
    for (int i = 1; i < M; ++i) {
      const float xi = xs[i * THREADS + threadIdx.x];
      for (int j = 0; j < i; ++j) pair += fabsf(xi - xs[j * THREADS + threadIdx.x]);
    }
    if (M > 1) spread = 2.0f * pair / (static_cast<float>(M) * static_cast<float>(M - 1));
    skill /= static_cast<float>(M);
    msum /= static_cast<float>(M);
  }
  if (!REDUCE) {
    if (active) {
      if (out_skill) out_skill[n * HW + p] = skill;
      if (out_spread) out_spread[n * HW + p] = spread;
      if (out_mean) out_mean[n * HW + p] = msum;
    }
    return;
  }
  // [se, skill, spread, crps] weighted sums + non-NaN counts
  double v[8];
  {
    const double w = active ? latw[p / W] : 0.0;
    const float d = msum - y;
    const double se = static_cast<double>(d * d) * w;
    const double sk = static_cast<double>(skill) * w;
    const double sp = static_cast<double>(spread) * w;
    const double cr = sk - 0.5 * sp;
    const double vals[4] = {se, sk, sp, cr};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const bool ok = active && !isnan(vals[k]);
      v[k] = ok ? vals[k] : 0.0;
      v[4 + k] = ok ? 1.0 : 0.0;
    }
  }
  __shared__ double red[8][THREADS / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const double s = warp_sum_d(v[k]);
    if (lane == 0) red[k][wid] = s;
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    double s = 0.0;
    for (int i = 0; i < THREADS / 32; ++i) s += red[threadIdx.x][i];
    if (threadIdx.x < 4) atomicAdd(&sums[threadIdx.x * N + n], s);
    else atomicAdd(&counts[(threadIdx.x - 4) * N + n], s);
  }
}

// Anomaly correlation coefficient terms (evaluate/utils.py:122-149): per plane, NaN-skipping weighted sums of
// fa*ta, fa^2, ta^2 (fa = forecast - climate, ta = truth - climate) and their non-NaN counts.
__global__ void __launch_bounds__(THREADS) acc_kernel(const float* __restrict__ fc, const float* __restrict__ tr,
                                                      const float* __restrict__ cl, const double* __restrict__ latw,
                                                      long long N, int H, int W, double* __restrict__ sums,
                                                      double* __restrict__ counts) {
  const int HW = H * W;
  const long long n = blockIdx.y;
  const int p = blockIdx.x * THREADS + threadIdx.x;
  double v[6] = {0, 0, 0, 0, 0, 0};
  if (p < HW) {
    const double w = latw != nullptr ? latw[p / W] : 1.0;
    const float c = cl[n * HW + p];
    const float fa = fc[n * HW + p] - c, ta = tr[n * HW + p] - c;
    const double vals[3] = {static_cast<double>(fa * ta) * w, static_cast<double>(fa * fa) * w,
                            static_cast<double>(ta * ta) * w};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const bool ok = !isnan(vals[k]);
      v[k] = ok ? vals[k] : 0.0;
      v[3 + k] = ok ? 1.0 : 0.0;
    }
  }
  __shared__ double red[6][THREADS / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const double s = warp_sum_d(v[k]);
    if (lane == 0) red[k][wid] = s;
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    double s = 0.0;
    for (int i = 0; i < THREADS / 32; ++i) s += red[threadIdx.x][i];
    if (threadIdx.x < 3) atomicAdd(&sums[threadIdx.x * N + n], s);
    else atomicAdd(&counts[(threadIdx.x - 3) * N + n], s);
  }
}

}  // namespace
}  // namespace lc

using namespace lc;

extern "C" {

int lc_metrics_accumulate(const float* fields, const float* truth, const double* latw, int members, long long planes,
                          int height, int width, double* sums, double* counts, void* stream) {
  LC_REQUIRE(fields && truth && latw && sums && counts, "null argument");
  LC_REQUIRE(members >= 1 && members <= 128, "ensemble size must be in [1, 128]");
  LC_REQUIRE(planes > 0 && planes <= 65535, "number of (channel, lead) planes must be in [1, 65535] per call");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t smem = static_cast<size_t>(members) * THREADS * sizeof(float);
  static size_t attr = 0;
  if (smem > 48 * 1024 && smem > attr) {
    LC_CHECK_CUDA(cudaFuncSetAttribute(metrics_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr = smem;
  }
  LC_CHECK_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 4 * planes, st));
  LC_CHECK_CUDA(cudaMemsetAsync(counts, 0, sizeof(double) * 4 * planes, st));
  dim3 grid(ceil_div(height * width, THREADS), static_cast<unsigned>(planes));
  metrics_kernel<true><<<grid, THREADS, smem, st>>>(fields, truth, latw, members, planes, height, width, sums, counts,
                                                    nullptr, nullptr, nullptr);
  LC_LAUNCH_CHECK();
  return 0;
}

int lc_metrics_pointwise(const float* fields, const float* truth, int members, long long planes, int height, int width,
                         float* out_skill, float* out_spread, float* out_mean, void* stream) {
  LC_REQUIRE(fields != nullptr, "null argument");
  LC_REQUIRE(out_skill == nullptr || truth != nullptr, "CRPS skill needs the truth tensor");
  LC_REQUIRE(members >= 1 && members <= 128, "ensemble size must be in [1, 128]");
  LC_REQUIRE(planes > 0 && planes <= 65535, "number of planes must be in [1, 65535] per call");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t smem = static_cast<size_t>(members) * THREADS * sizeof(float);
  static size_t attr = 0;
  if (smem > 48 * 1024 && smem > attr) {
    LC_CHECK_CUDA(cudaFuncSetAttribute(metrics_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr = smem;
  }
  dim3 grid(ceil_div(height * width, THREADS), static_cast<unsigned>(planes));
  metrics_kernel<false><<<grid, THREADS, smem, st>>>(fields, truth, nullptr, members, planes, height, width, nullptr,
                                                     nullptr, out_skill, out_spread, out_mean);
  LC_LAUNCH_CHECK();
  return 0;
}

int lc_metrics_acc(const float* forecast, const float* truth, const float* climate, const double* lat_weights,
                   long long planes, int height, int width, double* sums, double* counts, void* stream) {
  LC_REQUIRE(forecast && truth && climate && sums && counts, "null argument");
  LC_REQUIRE(planes > 0 && planes <= 65535, "number of planes must be in [1, 65535] per call");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LC_CHECK_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 3 * planes, st));
  LC_CHECK_CUDA(cudaMemsetAsync(counts, 0, sizeof(double) * 3 * planes, st));
  dim3 grid(ceil_div(height * width, THREADS), static_cast<unsigned>(planes));
  acc_kernel<<<grid, THREADS, 0, st>>>(forecast, truth, climate, lat_weights, planes, height, width, sums, counts);
  LC_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
