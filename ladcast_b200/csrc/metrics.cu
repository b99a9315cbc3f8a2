// Ensemble forecast metrics on the device: latitude-weighted ensemble-mean squared error, CRPS skill / spread /
// total.  Reference: evaluate/utils.py:40-118 (pointwise functions) and the assembly loop
// evaluate/evaluate_ens_gpu.py:339-415 (weights are float64, the SST channel is reduced with nanmean).
//
//   fields: M members x N (channel, lead) planes x HW pixels, f32; member m starts at fields + m * member_stride and
//   its planes are contiguous ([N, HW]) — a contiguous [M, C, T, H, W] tensor, or the per-member receive slots of
//   the multi-GPU exchange, are read in place.        truth [N, HW] f32 (NaN allowed)
//
// One thread per pixel (consecutive threads = consecutive pixels: every member load is a coalesced 128-B line per
// warp, all M loads of a thread in flight).  The M member values live in REGISTERS; the spread follows the
// reference's formulation exactly — sort the members (evaluate/utils.py:86, here a fully unrolled merge-exchange
// network for exactly M values) and form 2/(M(M-1)) * sum_i (2i - M - 1) x_(i) in fp32 (:91-99).  Per-pixel
// values are fp32 (as in the reference), the latitude-weighted spatial sums are fp64: per-thread accumulation over a
// few pixels, warp-shuffle + shared-memory block reduction, one fp64 atomicAdd per block and metric.
// Ensembles larger than 64 members use a shared-memory pairwise kernel (mean |x_i - x_j|, algebraically identical).
#include <cmath>

#include "../../include/ladcast_b200.h"
#include "common.cuh"

namespace lc {
namespace {

constexpr int THREADS = 256;
constexpr int PIX_PER_THREAD = 4;  // pixels a thread walks over (amortises the fp64 block reduction)

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sorting network for EXACTLY M values in registers: Knuth's merge exchange (TAOCP 5.2.2, Algorithm M = Batcher's
// odd-even merge sort for arbitrary n): 97 comparators for 20 members, 395 for 50 (a bitonic network padded to the next
// power of two needs 240 / 672).  The comparator list is built at compile time; after full unrolling every index is
// a constant, so the values never leave the register file.
template <int M>
struct MergeExchangeNet {
  static constexpr int kMax = M * 6 * 6 / 4 + 8;  // >= M * ceil(log2 M)^2 / 4 comparators
  int n = 0;
  int a[kMax] = {}, b[kMax] = {};
  constexpr MergeExchangeNet() {
    int t = 0;
    while ((1 << t) < M) ++t;
    if (M < 2) return;
    for (int p = 1 << (t - 1); p > 0; p >>= 1) {
      int q = 1 << (t - 1), r = 0, d = p;
      while (true) {
        for (int i = 0; i < M - d; ++i)
          if ((i & p) == r) { a[n] = i; b[n] = i + d; ++n; }
        if (q == p) break;
        d = q - p; q >>= 1; r = p;
      }
    }
  }
};

template <int M>
__device__ __forceinline__ void sort_network(float (&v)[M]) {
  constexpr MergeExchangeNet<M> net;
#pragma unroll
  for (int c = 0; c < net.n; ++c) {
    const float lo = fminf(v[net.a[c]], v[net.b[c]]), hi = fmaxf(v[net.a[c]], v[net.b[c]]);
    v[net.a[c]] = lo;
    v[net.b[c]] = hi;
  }
}

// block reduction of NV fp64 values per thread -> atomicAdd into out[k * stride + n]
template <int NV>
__device__ __forceinline__ void block_accumulate(const double (&v)[NV], double* __restrict__ first, double* __restrict__ second,
                                                 long long stride, long long n) {
  __shared__ double red[NV][THREADS / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const double s = warp_sum_d(v[k]);
    if (lane == 0) red[k][wid] = s;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = 0.0;
    for (int i = 0; i < THREADS / 32; ++i) s += red[threadIdx.x][i];
    constexpr int HALF = NV / 2;
    if (threadIdx.x < HALF) atomicAdd(&first[threadIdx.x * stride + n], s);
    else atomicAdd(&second[(threadIdx.x - HALF) * stride + n], s);
  }
}

__device__ __forceinline__ void add_weighted(double (&acc)[8], float msum, float skill, float spread, float y, double w) {
  // [se, skill, spread, crps] latitude-weighted values + non-NaN counts (nanmean of the SST channel)
  const float d = msum - y;
  const double se = static_cast<double>(d * d) * w;
  const double sk = static_cast<double>(skill) * w;
  const double sp = static_cast<double>(spread) * w;
  const double cr = sk - 0.5 * sp;
  const double vals[4] = {se, sk, sp, cr};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const bool ok = !isnan(vals[k]);
    acc[k] += ok ? vals[k] : 0.0;
    acc[4 + k] += ok ? 1.0 : 0.0;
  }
}

// member m of plane n, pixel p: (a) strided members of one allocation, (b) one base pointer per member — the members
// of other GPUs are read IN PLACE over NVLink (peer memory mapped through CUDA IPC), so the member->plane exchange and
// the reduction are one kernel and the gathered copy never exists.
struct StridedMembers {
  const float* base;
  long long member_stride;
  __device__ __forceinline__ float operator()(int m, long long off) const { return __ldg(base + m * member_stride + off); }
};
constexpr int MAX_PTR_MEMBERS = 64;
struct MemberPtrs {
  const float* p[MAX_PTR_MEMBERS];
};

template <int M, bool REDUCE, typename Load>
__device__ __forceinline__ void metrics_sorted_body(const Load& load, const float* __restrict__ truth,
                                                    const double* __restrict__ latw, long long N, int H, int W, int bpp,
                                                    double* __restrict__ sums, double* __restrict__ counts,
                                                    float* __restrict__ out_skill, float* __restrict__ out_spread,
                                                    float* __restrict__ out_mean) {
  const int HW = H * W;
  const long long n = blockIdx.x / bpp;
  const int blk = static_cast<int>(blockIdx.x - n * bpp);
  const float inv_m = 1.0f / static_cast<float>(M);
  const float spread_scale = M > 1 ? 2.0f / (static_cast<float>(M) * static_cast<float>(M - 1)) : 0.f;
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int p = blk * THREADS + threadIdx.x; p < HW; p += bpp * THREADS) {
    float x[M];
#pragma unroll
    for (int m = 0; m < M; ++m) x[m] = load(m, n * HW + p);
    const float y = truth != nullptr ? __ldg(truth + n * HW + p) : 0.f;
    float msum = 0.f, skill = 0.f;
#pragma unroll
    for (int m = 0; m < M; ++m) {
      msum += x[m];
      skill += fabsf(y - x[m]);
    }
    sort_network<M>(x);
    float ws = 0.f;
#pragma unroll
    for (int m = 0; m < M; ++m) ws = fmaf(static_cast<float>(2 * (m + 1) - M - 1), x[m], ws);
    float spread = spread_scale * ws;
    if (msum != msum) spread = msum;  // a NaN member: torch.sort keeps it and the weighted sum propagates it
    skill *= inv_m;
    msum *= inv_m;
    if (REDUCE) {
      add_weighted(acc, msum, skill, spread, y, latw[p / W]);
    } else {
      if (out_skill) out_skill[n * HW + p] = skill;
      if (out_spread) out_spread[n * HW + p] = spread;
      if (out_mean) out_mean[n * HW + p] = msum;
    }
  }
  if (REDUCE) block_accumulate<8>(acc, sums, counts, N, n);
}

template <int M, bool REDUCE>
__global__ void __launch_bounds__(THREADS) metrics_sorted_kernel(const float* __restrict__ fields, long long member_stride,
                                                                 const float* __restrict__ truth,
                                                                 const double* __restrict__ latw, long long N, int H,
                                                                 int W, int bpp, double* __restrict__ sums,
                                                                 double* __restrict__ counts, float* __restrict__ out_skill,
                                                                 float* __restrict__ out_spread, float* __restrict__ out_mean) {
  pdl_grid_sync();
  metrics_sorted_body<M, REDUCE>(StridedMembers{fields, member_stride}, truth, latw, N, H, W, bpp, sums, counts, out_skill,
                                 out_spread, out_mean);
}

// members given by pointer (local or peer memory): member m's planes are contiguous [N, HW] at ptrs.p[m]
template <int M>
__global__ void __launch_bounds__(THREADS) metrics_sorted_ptr_kernel(const __grid_constant__ MemberPtrs ptrs,
                                                                     const float* __restrict__ truth,
                                                                     const double* __restrict__ latw, long long N, int H,
                                                                     int W, int bpp, double* __restrict__ sums,
                                                                     double* __restrict__ counts) {
  pdl_grid_sync();
  auto load = [&ptrs](int m, long long off) { return ptrs.p[m][off]; };
  metrics_sorted_body<M, true>(load, truth, latw, N, H, W, bpp, sums, counts, nullptr, nullptr, nullptr);
}

// M > 64: members staged in shared memory (one column per thread), spread = mean absolute pair difference
template <bool REDUCE>
__global__ void __launch_bounds__(THREADS) metrics_pairwise_kernel(const float* __restrict__ fields, long long member_stride,
                                                                   const float* __restrict__ truth,
                                                                   const double* __restrict__ latw, int M, long long N, int H,
                                                                   int W, int bpp, double* __restrict__ sums,
                                                                   double* __restrict__ counts, float* __restrict__ out_skill,
                                                                   float* __restrict__ out_spread, float* __restrict__ out_mean) {
  pdl_grid_sync();
  extern __shared__ float xs[];  // [M][THREADS]
  const int HW = H * W;
  const long long n = blockIdx.x / bpp;
  const int blk = static_cast<int>(blockIdx.x - n * bpp);
  const float* base = fields + n * HW;
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int p = blk * THREADS + threadIdx.x; p < HW; p += bpp * THREADS) {
    const float y = truth != nullptr ? truth[n * HW + p] : 0.f;
    float msum = 0.f, skill = 0.f, spread = 0.f;
    for (int m = 0; m < M; ++m) {
      const float x = base[m * member_stride + p];
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
    if (REDUCE) {
      add_weighted(acc, msum, skill, spread, y, latw[p / W]);
    } else {
      if (out_skill) out_skill[n * HW + p] = skill;
      if (out_spread) out_spread[n * HW + p] = spread;
      if (out_mean) out_mean[n * HW + p] = msum;
    }
  }
  if (REDUCE) block_accumulate<8>(acc, sums, counts, N, n);
}

// Anomaly correlation coefficient terms (evaluate/utils.py:122-149): per plane, NaN-skipping weighted sums of
// fa*ta, fa^2, ta^2 (fa = forecast - climate, ta = truth - climate) and their non-NaN counts.
__global__ void __launch_bounds__(THREADS) acc_kernel(const float* __restrict__ fc, const float* __restrict__ tr,
                                                      const float* __restrict__ cl, const double* __restrict__ latw,
                                                      long long N, int H, int W, int bpp, double* __restrict__ sums,
                                                      double* __restrict__ counts) {
  pdl_grid_sync();
  const int HW = H * W;
  const long long n = blockIdx.x / bpp;
  const int blk = static_cast<int>(blockIdx.x - n * bpp);
  double v[6] = {0, 0, 0, 0, 0, 0};
  for (int p = blk * THREADS + threadIdx.x; p < HW; p += bpp * THREADS) {
    const double w = latw != nullptr ? latw[p / W] : 1.0;
    const float c = cl[n * HW + p];
    const float fa = fc[n * HW + p] - c, ta = tr[n * HW + p] - c;
    const double vals[3] = {static_cast<double>(fa * ta) * w, static_cast<double>(fa * fa) * w,
                            static_cast<double>(ta * ta) * w};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const bool ok = !isnan(vals[k]);
      v[k] += ok ? vals[k] : 0.0;
      v[3 + k] += ok ? 1.0 : 0.0;
    }
  }
  block_accumulate<6>(v, sums, counts, N, n);
}

int blocks_per_plane(int HW) { return ceil_div(HW, THREADS * PIX_PER_THREAD); }

template <bool REDUCE>
int launch_metrics(const float* fields, long long member_stride, const float* truth, const double* latw, int M, long long N,
                   int H, int W, double* sums, double* counts, float* o_skill, float* o_spread, float* o_mean,
                   cudaStream_t st) {
  const int bpp = blocks_per_plane(H * W);
  LC_REQUIRE(N * bpp < (1ll << 31), "too many (channel, lead) planes for one launch");
  const unsigned grid = static_cast<unsigned>(N * bpp);
  // algorithmic bytes: every member value and the truth read once (+ per-pixel outputs of the pointwise variant)
  const double px = static_cast<double>(N) * H * W;
  ProfScope ps(PROF_METRICS, 0.0, px * 4.0 * (M + (truth ? 1 : 0) + (o_skill ? 1 : 0) + (o_spread ? 1 : 0) + (o_mean ? 1 : 0)), st);
#define LC_SORTED(MM)                                                                                                   \
  case MM:                                                                                                              \
    LC_CHECK_CUDA(launch_kernel(metrics_sorted_kernel<MM, REDUCE>, grid, THREADS, 0, st, fields, member_stride, truth, latw, N, H, W, bpp, sums, \
                                                                counts, o_skill, o_spread, o_mean));                     \
    break;
#define LC_SORTED8(B) LC_SORTED(B) LC_SORTED(B + 1) LC_SORTED(B + 2) LC_SORTED(B + 3) LC_SORTED(B + 4) LC_SORTED(B + 5) \
    LC_SORTED(B + 6) LC_SORTED(B + 7)
  if (M <= 64) {
    switch (M) {
      LC_SORTED8(1) LC_SORTED8(9) LC_SORTED8(17) LC_SORTED8(25) LC_SORTED8(33) LC_SORTED8(41) LC_SORTED8(49) LC_SORTED8(57)
    }
  }
  else {
    const size_t smem = static_cast<size_t>(M) * THREADS * sizeof(float);
    static PerDevice<size_t> attr;
    if (smem > 48 * 1024 && smem > attr.here()) {
      LC_CHECK_CUDA(cudaFuncSetAttribute(metrics_pairwise_kernel<REDUCE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem)));
      attr.here() = smem;
    }
    LC_CHECK_CUDA(launch_kernel(metrics_pairwise_kernel<REDUCE>, grid, THREADS, smem, st, fields, member_stride, truth, latw, M, N, H, W, bpp, sums,
                                                                 counts, o_skill, o_spread, o_mean));
  }
#undef LC_SORTED8
#undef LC_SORTED
  LC_LAUNCH_CHECK();
  return 0;
}

int launch_metrics_ptrs(const MemberPtrs& ptrs, const float* truth, const double* latw, int M, long long N, int H, int W,
                        double* sums, double* counts, cudaStream_t st) {
  const int bpp = blocks_per_plane(H * W);
  LC_REQUIRE(N * bpp < (1ll << 31), "too many (channel, lead) planes for one launch");
  const unsigned grid = static_cast<unsigned>(N * bpp);
  ProfScope ps(PROF_METRICS, 0.0, static_cast<double>(N) * H * W * 4.0 * (M + 1), st);
#define LC_SORTED(MM)                                                                                                       \
  case MM:                                                                                                                  \
    LC_CHECK_CUDA(launch_kernel(metrics_sorted_ptr_kernel<MM>, grid, THREADS, 0, st, ptrs, truth, latw, N, H, W, bpp, sums, counts)); \
    break;
#define LC_SORTED8(B) LC_SORTED(B) LC_SORTED(B + 1) LC_SORTED(B + 2) LC_SORTED(B + 3) LC_SORTED(B + 4) LC_SORTED(B + 5) \
    LC_SORTED(B + 6) LC_SORTED(B + 7)
  switch (M) {
    LC_SORTED8(1) LC_SORTED8(9) LC_SORTED8(17) LC_SORTED8(25) LC_SORTED8(33) LC_SORTED8(41) LC_SORTED8(49) LC_SORTED8(57)
  }
#undef LC_SORTED8
#undef LC_SORTED
  LC_LAUNCH_CHECK();
  return 0;
}

}  // namespace
}  // namespace lc

using namespace lc;

extern "C" {

int lc_metrics_accumulate_strided(const float* fields, long long member_stride, const float* truth, const double* latw,
                                  int members, long long planes, int height, int width, double* sums, double* counts,
                                  void* stream) {
  LC_REQUIRE(fields && truth && latw && sums && counts, "null argument");
  LC_REQUIRE(members >= 1 && members <= 128, "ensemble size must be in [1, 128]");
  LC_REQUIRE(planes > 0 && height > 0 && width > 0, "bad shape");
  LC_REQUIRE(member_stride >= planes * height * width || members == 1, "member stride smaller than one member's planes");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LC_CHECK_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 4 * planes, st));
  LC_CHECK_CUDA(cudaMemsetAsync(counts, 0, sizeof(double) * 4 * planes, st));
  return launch_metrics<true>(fields, member_stride, truth, latw, members, planes, height, width, sums, counts, nullptr,
                              nullptr, nullptr, st);
}

int lc_metrics_accumulate(const float* fields, const float* truth, const double* latw, int members, long long planes,
                          int height, int width, double* sums, double* counts, void* stream) {
  return lc_metrics_accumulate_strided(fields, planes * height * width, truth, latw, members, planes, height, width, sums,
                                       counts, stream);
}

int lc_metrics_accumulate_ptrs(const float* const* member_ptrs, const float* truth, const double* latw, int members,
                               long long planes, int height, int width, double* sums, double* counts, void* stream) {
  LC_REQUIRE(member_ptrs && truth && latw && sums && counts, "null argument");
  LC_REQUIRE(members >= 1 && members <= MAX_PTR_MEMBERS, "ensemble size must be in [1, 64] for the pointer form");
  LC_REQUIRE(planes > 0 && height > 0 && width > 0, "bad shape");
  MemberPtrs ptrs;
  for (int m = 0; m < MAX_PTR_MEMBERS; ++m) ptrs.p[m] = m < members ? member_ptrs[m] : nullptr;
  for (int m = 0; m < members; ++m) LC_REQUIRE(ptrs.p[m] != nullptr, "null member pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LC_CHECK_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 4 * planes, st));
  LC_CHECK_CUDA(cudaMemsetAsync(counts, 0, sizeof(double) * 4 * planes, st));
  return launch_metrics_ptrs(ptrs, truth, latw, members, planes, height, width, sums, counts, st);
}

int lc_metrics_pointwise(const float* fields, const float* truth, int members, long long planes, int height, int width,
                         float* out_skill, float* out_spread, float* out_mean, void* stream) {
  LC_REQUIRE(fields != nullptr, "null argument");
  LC_REQUIRE(out_skill == nullptr || truth != nullptr, "CRPS skill needs the truth tensor");
  LC_REQUIRE(members >= 1 && members <= 128, "ensemble size must be in [1, 128]");
  LC_REQUIRE(planes > 0 && height > 0 && width > 0, "bad shape");
  return launch_metrics<false>(fields, planes * height * width, truth, nullptr, members, planes, height, width, nullptr,
                               nullptr, out_skill, out_spread, out_mean, static_cast<cudaStream_t>(stream));
}

int lc_metrics_acc(const float* forecast, const float* truth, const float* climate, const double* lat_weights,
                   long long planes, int height, int width, double* sums, double* counts, void* stream) {
  LC_REQUIRE(forecast && truth && climate && sums && counts, "null argument");
  LC_REQUIRE(planes > 0 && height > 0 && width > 0, "bad shape");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LC_CHECK_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 3 * planes, st));
  LC_CHECK_CUDA(cudaMemsetAsync(counts, 0, sizeof(double) * 3 * planes, st));
  const int bpp = blocks_per_plane(height * width);
  LC_REQUIRE(planes * bpp < (1ll << 31), "too many planes for one launch");
  ProfScope ps(PROF_METRICS, 0.0, static_cast<double>(planes) * height * width * 12.0, st);
  LC_CHECK_CUDA(launch_kernel(acc_kernel, static_cast<unsigned>(planes * bpp), THREADS, 0, st, forecast, truth, climate, lat_weights, planes, height,
                                                                     width, bpp, sums, counts));
  LC_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
