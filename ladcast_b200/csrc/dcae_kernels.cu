// HBM-bound kernels of the DC-AE decoder (NHWC internal layout): sphere padding, depthwise sphere convolutions,
// grouped 1x1, ReLU linear attention, channel RMSNorm (+residual), pixel-shuffle (+shortcut).
// Reference: models/sphere_conv.py:62-192, models/DCAE.py:96-324, 327-377, 493-536, 717-732.
#include "dcae_kernels.h"

namespace lc {
namespace {

// Source pixel of padded position (py, px) for a sphere pad of p rows/cols (sphere_conv.py:62-91):
// longitude circular; pole rows = first/last p rows flipped vertically and rolled by W/2.
__device__ __forceinline__ void sphere_src(int py, int px, int p, int H, int W, int& sy, int& sx) {
  int cx = px - p;
  cx = (cx % W + W) % W;
  if (py < p) {
    sy = p - 1 - py;
    sx = (cx - W / 2 + W) % W;
  } else if (py >= H + p) {
    sy = H - 1 - (py - (H + p));
    sx = (cx - W / 2 + W) % W;
  } else {
    sy = py - p;
    sx = cx;
  }
}

template <typename T>
__device__ __forceinline__ float ld(const T* p) { return to_f32<T>(*p); }

__device__ __forceinline__ float plane_load(const PlaneSrc& z, int f, int c, int pix) {
  float v = z.z[z.plane(f, c) + pix];
  if (z.scale != nullptr) v = __fadd_rn(__fmul_rn(__fdiv_rn(v, z.target), __ldg(z.scale + c)), __ldg(z.shift + c));
  return v;
}

// ---------------------------------------------------------------- NCHW f32 latent -> padded NHWC T  (conv_in input)
template <typename T>
__global__ void pad_from_nchw_kernel(PlaneSrc z, T* __restrict__ out, int n, int C, int H, int W, int Cp) {
  pdl_grid_sync();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(n) * (H + 2) * (W + 2) * Cp;
  if (i >= total) return;
  const int c = static_cast<int>(i % Cp);
  long long r = i / Cp;
  const int px = static_cast<int>(r % (W + 2));
  r /= (W + 2);
  const int py = static_cast<int>(r % (H + 2));
  const int f = static_cast<int>(r / (H + 2));
  float v = 0.f;
  if (c < C) {
    int sy, sx;
    sphere_src(py, px, 1, H, W, sy, sx);
    v = plane_load(z, f, c, sy * W + sx);
  }
  out[i] = from_f32<T>(v);
}

// ---------------------------------------------------------------- NHWC f32 [n,H,W,C] -> padded NHWC T [n,H+2,W+2,Cp]
template <typename T>
__global__ void pad_from_nhwc_kernel(const float* __restrict__ x, T* __restrict__ out, int n, int C, int H, int W,
                                     int Cp) {
  pdl_grid_sync();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;  // one thread per 4 channels
  const int c4 = Cp / 4;
  const long long total = static_cast<long long>(n) * (H + 2) * (W + 2) * c4;
  if (i >= total) return;
  const int c = static_cast<int>(i % c4) * 4;
  long long r = i / c4;
  const int px = static_cast<int>(r % (W + 2));
  r /= (W + 2);
  const int py = static_cast<int>(r % (H + 2));
  const int f = static_cast<int>(r / (H + 2));
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < C) {
    int sy, sx;
    sphere_src(py, px, 1, H, W, sy, sx);
    v = *reinterpret_cast<const float4*>(x + ((static_cast<long long>(f) * H + sy) * W + sx) * C + c);
  }
  T* o = out + i * 4;
  o[0] = from_f32<T>(v.x); o[1] = from_f32<T>(v.y); o[2] = from_f32<T>(v.z); o[3] = from_f32<T>(v.w);
}

// ---------------------------------------------------------------- fill the 1-pixel halo of a padded NHWC buffer
// from its own interior (after a conv epilogue wrote the interior directly).
template <typename T>
__global__ void halo_fill_kernel(T* __restrict__ buf, int n, int H, int W, int Cp) {
  pdl_grid_sync();
  // halo positions per frame: 2 full rows (W+2) + 2 columns x H
  const int per_frame = 2 * (W + 2) + 2 * H;
  const int c8 = Cp / 8;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(n) * per_frame * c8;
  if (i >= total) return;
  const int c = static_cast<int>(i % c8) * 8;
  long long r = i / c8;
  const int hidx = static_cast<int>(r % per_frame);
  const int f = static_cast<int>(r / per_frame);
  int py, px;
  if (hidx < W + 2) { py = 0; px = hidx; }
  else if (hidx < 2 * (W + 2)) { py = H + 1; px = hidx - (W + 2); }
  else { const int k = hidx - 2 * (W + 2); py = 1 + (k >> 1); px = (k & 1) ? W + 1 : 0; }
  int sy, sx;
  sphere_src(py, px, 1, H, W, sy, sx);
  const long long fs = static_cast<long long>(f) * (H + 2) * (W + 2);
  const T* src = buf + (fs + static_cast<long long>(sy + 1) * (W + 2) + (sx + 1)) * Cp + c;
  T* dst = buf + (fs + static_cast<long long>(py) * (W + 2) + px) * Cp + c;
  if (sizeof(T) == 2) {
    *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);
  } else {
    reinterpret_cast<float4*>(dst)[0] = reinterpret_cast<const float4*>(src)[0];
    reinterpret_cast<float4*>(dst)[1] = reinterpret_cast<const float4*>(src)[1];
  }
}

// ---------------------------------------------------------------- depthwise KxK sphere conv on NHWC (unpadded input)
// out[p, c] = bias[c] + sum_{ky,kx} w[ky, kx', c] * in[src(y+ky, x+kx), c], kx' mirrored in the pole pad rows of
// output rows 0 / H-1 (sphere_conv.py:93-129).  Weights are tap-major [K*K, C] so channel vectors load coalesced.
__device__ __forceinline__ void sphere_row(int py, int p, int H, int& sy, bool& rolled) {
  if (py < p) { sy = p - 1 - py; rolled = true; }
  else if (py >= H + p) { sy = H - 1 - (py - (H + p)); rolled = true; }
  else { sy = py - p; rolled = false; }
}
// Depthwise kernels: lane = channel quad (a warp reads 128 consecutive channels of one pixel: whole 128-B lines),
// each thread slides along XT consecutive x so every loaded input vector feeds up to K outputs and the K taps of a
// kernel row are loaded once per XT outputs.  The kernels are L1-wavefront bound, not DRAM bound (profiles/): this
// cuts the wavefronts per output ~3x against one-output-per-thread.
__device__ __forceinline__ int sphere_col0(int x0, int p, int W, bool rolled) {
  int cx = (x0 - p + (rolled ? W / 2 : 0)) % W;
  return cx < 0 ? cx + W : cx;
}

// a += w * v on the packed fp32 pipe (FFMA2: two lanes per issue slot; these kernels are issue-bound)
__device__ __forceinline__ void fma4(float4& a, const float4& w, const float4& v) {
  const float2 lo = __ffma2_rn(make_float2(w.x, w.y), make_float2(v.x, v.y), make_float2(a.x, a.y));
  const float2 hi = __ffma2_rn(make_float2(w.z, w.w), make_float2(v.z, v.w), make_float2(a.z, a.w));
  a = make_float4(lo.x, lo.y, hi.x, hi.y);
}
// element offsets of the NJ source columns a thread slides over (x0-p .. x0-p+NJ-1, wrapped; +W/2 on pole rows)
template <int NJ>
__device__ __forceinline__ void col_offsets(int (&co)[NJ], int x0, int p, int W, int C, bool rolled) {
  int cx = sphere_col0(x0, p, W, rolled);
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    co[j] = cx * C;
    if (++cx == W) cx = 0;
  }
}
__device__ __forceinline__ float silu_fast(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

// 3x3 + bias + GLU: channels [value | gate] halves -> value * silu(gate); T in/out
// (GLUMBConv.conv_depth + chunk + nonlinearity, DCAE.py:312-315)
template <typename T>
__device__ __forceinline__ float4 ld4(const T* p);
template <>
__device__ __forceinline__ float4 ld4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <>
__device__ __forceinline__ float4 ld4<bf16>(const bf16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
template <typename T>
__device__ __forceinline__ void st4(T* p, float4 v);
template <>
__device__ __forceinline__ void st4<float>(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
template <>
__device__ __forceinline__ void st4<bf16>(bf16* p, float4 v) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&lo);
  u.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(p) = u;
}

template <typename T, int XT>
__global__ void __launch_bounds__(256, 3) dwconv3_glu_kernel(const T* __restrict__ in, const float* __restrict__ w,
                                                          const float* __restrict__ bias, T* __restrict__ out, int n,
                                                          int H, int W, int C) {
  pdl_grid_sync();
  const int Co = C / 2;
  const int c = (blockIdx.x * 32 + (threadIdx.x & 31)) * 4;
  const int x0 = (blockIdx.y * 8 + (threadIdx.x >> 5)) * XT;
  const int y = blockIdx.z % H, f = blockIdx.z / H;
  if (c >= Co || x0 >= W) return;
  const T* base = in + static_cast<long long>(f) * H * W * C + c;
  float4 a0[XT], a1[XT];
  {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + Co + c));
#pragma unroll
    for (int o = 0; o < XT; ++o) { a0[o] = b0; a1[o] = b1; }
  }
  int con[XT + 2];
  col_offsets<XT + 2>(con, x0, 1, W, C, false);
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    int sy;
    bool rolled;
    sphere_row(y + ky, 1, H, sy, rolled);
    const bool flip = (y == 0 && ky == 0) || (y == H - 1 && ky == 2);
    float4 wv[3], wg[3];
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const float* wt = w + (ky * 3 + (flip ? 2 - kx : kx)) * C + c;
      wv[kx] = __ldg(reinterpret_cast<const float4*>(wt));
      wg[kx] = __ldg(reinterpret_cast<const float4*>(wt + Co));
    }
    const T* rowp = base + static_cast<long long>(sy) * W * C;
    int co[XT + 2];
#pragma unroll
    for (int j = 0; j < XT + 2; ++j) co[j] = con[j];
    if (rolled) col_offsets<XT + 2>(co, x0, 1, W, C, true);
#pragma unroll
    for (int j = 0; j < XT + 2; ++j) {
      const T* px = rowp + co[j];
      const float4 v = ld4<T>(px), g = ld4<T>(px + Co);
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int o = j - kx;
        if (o >= 0 && o < XT) { fma4(a0[o], wv[kx], v); fma4(a1[o], wg[kx], g); }
      }
    }
  }
  T* op = out + ((static_cast<long long>(f) * H + y) * W + x0) * Co + c;
#pragma unroll
  for (int o = 0; o < XT; ++o)
    if (x0 + o < W)
      st4<T>(op + static_cast<long long>(o) * Co, make_float4(a0[o].x * silu_fast(a1[o].x), a0[o].y * silu_fast(a1[o].y),
                                                             a0[o].z * silu_fast(a1[o].z), a0[o].w * silu_fast(a1[o].w)));
}

// ---------------------------------------------------------------- fused multiscale projection (DCAE.py:76-88)
// depthwise 5x5 sphere conv -> grouped 32->32 1x1 conv in one pass; the intermediate never leaves the SM.
// Block = 128 channels (4 groups) x one image row, walked in 64-pixel segments:
//   phase A: sliding-window 5x5 (lane = channel quad, 4 outputs per thread), results stored channel-major into ds[ch][px] (XOR-swizzled 16-B chunks so
//            both the transposing stores and the phase-B reads are bank-conflict free);
//   phase B: warp = (group, 32-pixel half); thread tile 8 px x 4 outputs, outer product over the 32 inputs of the
//            group: 3 LDS.128 per 32 FFMA (the stand-alone kernel's broadcast reads cost 8 per 32).
template <typename T>
__global__ void __launch_bounds__(256, 3) multiscale_fused_kernel(const T* __restrict__ in, const float* __restrict__ w5,
                                                               const float* __restrict__ wg, T* __restrict__ out,
                                                               int n, int H, int W, int C) {
  pdl_grid_sync();
  __shared__ __align__(16) float ds[128][64];
  __shared__ __align__(16) float ws[4][32][32];  // [group][k][o ^ swizzle(k)]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ncb = (C + 127) / 128;
  const int cb = (blockIdx.x % ncb) * 128;
  const int fy = blockIdx.x / ncb;
  const int y = fy % H, f = fy / H;
  // group weights, transposed to k-major
  for (int i = tid; i < 4096; i += 256) {
    const int g = i >> 10, o = (i >> 5) & 31, k = i & 31;
    float v = 0.f;
    if (cb + g * 32 < C) v = wg[(static_cast<long long>(cb) + g * 32 + o) * 32 + k];
    ws[g][k][o ^ ((k & 7) * 4)] = v;
  }
  const int c = cb + lane * 4;
  const bool c_ok = c < C;
  const T* base = in + static_cast<long long>(f) * H * W * C + c;
  const int g = warp & 3, half = warp >> 2;
  const int nl = lane & 7, ml = lane >> 3;
  const bool g_ok = cb + g * 32 < C;
  for (int seg = 0; seg < W; seg += 64) {
    __syncthreads();  // ws ready / previous segment's phase B done
    // ---- phase A
#pragma unroll 1
    for (int it = 0; it < 2; ++it) {
      const int xl = (it * 8 + warp) * 4;
      const int x0 = seg + xl;
      if (!c_ok || x0 >= W) continue;
      float4 acc[4];
#pragma unroll
      for (int o = 0; o < 4; ++o) acc[o] = make_float4(0.f, 0.f, 0.f, 0.f);
      int con[8];
      col_offsets<8>(con, x0, 2, W, C, false);
#pragma unroll
      for (int ky = 0; ky < 5; ++ky) {
        int sy;
        bool rolled;
        sphere_row(y + ky, 2, H, sy, rolled);
        const bool flip = (y == 0 && ky < 2) || (y == H - 1 && ky >= 3);
        float4 wr[5];
#pragma unroll
        for (int kx = 0; kx < 5; ++kx)
          wr[kx] = __ldg(reinterpret_cast<const float4*>(w5 + (ky * 5 + (flip ? 4 - kx : kx)) * C + c));
        const T* rowp = base + static_cast<long long>(sy) * W * C;
        int co[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) co[j] = con[j];
        if (rolled) col_offsets<8>(co, x0, 2, W, C, true);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 v = ld4<T>(rowp + co[j]);
#pragma unroll
          for (int kx = 0; kx < 5; ++kx) {
            const int o = j - kx;
            if (o >= 0 && o < 4) fma4(acc[o], wr[kx], v);
          }
        }
      }
      const int sw = (lane & 7) * 4;  // swizzle of rows 4*lane .. 4*lane+3
      *reinterpret_cast<float4*>(&ds[lane * 4 + 0][xl ^ sw]) = make_float4(acc[0].x, acc[1].x, acc[2].x, acc[3].x);
      *reinterpret_cast<float4*>(&ds[lane * 4 + 1][xl ^ sw]) = make_float4(acc[0].y, acc[1].y, acc[2].y, acc[3].y);
      *reinterpret_cast<float4*>(&ds[lane * 4 + 2][xl ^ sw]) = make_float4(acc[0].z, acc[1].z, acc[2].z, acc[3].z);
      *reinterpret_cast<float4*>(&ds[lane * 4 + 3][xl ^ sw]) = make_float4(acc[0].w, acc[1].w, acc[2].w, acc[3].w);
    }
    __syncthreads();
    // ---- phase B
    const int px0 = half * 32 + ml * 8;
    if (!g_ok || seg + half * 32 >= W) continue;
    float2 a[4][4];  // [pixel pair][output]: packed FMA lanes = two adjacent pixels
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) a[i][j] = make_float2(0.f, 0.f);
#pragma unroll 4
    for (int k = 0; k < 32; ++k) {
      const int row = g * 32 + k;
      const int sw = ((row >> 2) & 7) * 4;
      const float4 xa = *reinterpret_cast<const float4*>(&ds[row][px0 ^ sw]);
      const float4 xb = *reinterpret_cast<const float4*>(&ds[row][(px0 + 4) ^ sw]);
      const float4 wv = *reinterpret_cast<const float4*>(&ws[g][k][(nl * 4) ^ ((k & 7) * 4)]);
      const float2 xs[4] = {make_float2(xa.x, xa.y), make_float2(xa.z, xa.w), make_float2(xb.x, xb.y),
                            make_float2(xb.z, xb.w)};
      const float2 w2[4] = {make_float2(wv.x, wv.x), make_float2(wv.y, wv.y), make_float2(wv.z, wv.z),
                            make_float2(wv.w, wv.w)};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) a[i][j] = __ffma2_rn(xs[i], w2[j], a[i][j]);
    }
    T* op = out + ((static_cast<long long>(f) * H + y) * W + seg + px0) * C + cb + g * 32 + nl * 4;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (seg + px0 + i < W)
        st4<T>(op + static_cast<long long>(i) * C,
               (i & 1) ? make_float4(a[i >> 1][0].y, a[i >> 1][1].y, a[i >> 1][2].y, a[i >> 1][3].y)
                       : make_float4(a[i >> 1][0].x, a[i >> 1][1].x, a[i >> 1][2].x, a[i >> 1][3].x));
  }
}

// ---------------------------------------------------------------- ReLU linear attention (DCAE.py:155-175, 226-262)
// One CTA per (frame, head).  Head g of scale s reads channels [g*96, g*96+96) of qkv (s=0) or of the multiscale
// branch (s=1): 32 q | 32 k | 32 v.  S = [V;1] relu(K)^T (33x32, fp32), out = S relu(Q), out[:32] / (out[32] + eps).
// Both phases are register-tiled (the kernel used to be bound by shared-memory wavefronts):
//   phase 1: each warp owns every 8th pixel and a private 33x32 partial of S; lane = (4 v-channels) x (8 k-channels),
//            operands straight from global memory as 128-B row segments, 4 pixels in flight, packed FFMA2;
//   phase 2: lane = (4 output channels) x (8-wide slice of the q dot product); 4 pixels per step, the partial dot
//            products are combined with a transposing butterfly (15 shuffles per 4 pixels) that leaves lane kl with
//            the finished sums of pixel kl.
__device__ __forceinline__ float2 relu2(float a, float b) { return make_float2(fmaxf(a, 0.f), fmaxf(b, 0.f)); }

template <typename TI, typename TO>
__global__ void __launch_bounds__(256, 2) linear_attn_kernel(const TI* __restrict__ qkv, const TI* __restrict__ ms,
                                                             TO* __restrict__ out, int HW, int heads, float eps) {
  pdl_grid_sync();
  __shared__ __align__(16) float red[8][33][32];  // per-warp partials of S
  __shared__ __align__(16) float S[33][32];
  const int g = blockIdx.x % (2 * heads);
  const int f = blockIdx.x / (2 * heads);
  const int scale = g / heads, hg = g % heads;
  const int C3 = heads * 96;
  const TI* src = (scale ? ms : qkv) + static_cast<long long>(f) * HW * C3 + hg * 96;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cl = lane >> 2, kl = lane & 3;

  // ---- phase 1: S[c][c'] += v[p][c] * relu(k[p][c']),  S[32][c'] += relu(k[p][c'])
  {
    float2 acc[4][4];  // [v channel 4cl+i][k channel pair 8kl+2j, +1]
    float2 ks[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ks[i] = make_float2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);
    }
    constexpr int U = 4;
    for (int p0 = warp * U; p0 < HW; p0 += 8 * U) {
      float4 v4[U], ka[U], kb[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        v4[u] = ka[u] = kb[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p0 + u < HW) {
          const TI* row = src + static_cast<long long>(p0 + u) * C3;
          v4[u] = ld4<TI>(row + 64 + cl * 4);
          ka[u] = ld4<TI>(row + 32 + kl * 8);
          kb[u] = ld4<TI>(row + 32 + kl * 8 + 4);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float2 k2[4] = {relu2(ka[u].x, ka[u].y), relu2(ka[u].z, ka[u].w), relu2(kb[u].x, kb[u].y),
                              relu2(kb[u].z, kb[u].w)};
        const float vv[4] = {v4[u].x, v4[u].y, v4[u].z, v4[u].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 vd = make_float2(vv[i], vv[i]);
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = __ffma2_rn(vd, k2[j], acc[i][j]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) ks[j] = __fadd2_rn(ks[j], k2[j]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) *reinterpret_cast<float2*>(&red[warp][cl * 4 + i][kl * 8 + 2 * j]) = acc[i][j];
    if (cl == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) *reinterpret_cast<float2*>(&red[warp][32][kl * 8 + 2 * j]) = ks[j];
    }
  }
  __syncthreads();
  for (int i = tid; i < 33 * 32; i += 256) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) a += (&red[w][0][0])[i];
    (&S[0][0])[i] = a;
  }
  __syncthreads();

  // ---- phase 2: out[p][c] = (S[c] . relu(q[p])) / (S[32] . relu(q[p]) + eps)
  float2 sr[5][4];  // rows 4cl..4cl+3 of S and the row of ones (index 4), k slice 8kl..8kl+7 as pairs
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    const float* srow = i < 4 ? &S[cl * 4 + i][kl * 8] : &S[32][kl * 8];
#pragma unroll
    for (int j = 0; j < 4; ++j) sr[i][j] = *reinterpret_cast<const float2*>(srow + 2 * j);
  }
  const int Co = 2 * heads * 32;
  const int my_px = (kl & 1) * 2 + (kl >> 1);  // pixel of the group of 4 this lane finishes
  for (int p0 = warp * 4; p0 < HW; p0 += 32) {
    float4 qa[4], qb[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      qa[u] = qb[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p0 + u < HW) {
        const TI* row = src + static_cast<long long>(p0 + u) * C3 + kl * 8;
        qa[u] = ld4<TI>(row);
        qb[u] = ld4<TI>(row + 4);
      }
    }
    float val[4][5];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float2 q2[4] = {relu2(qa[u].x, qa[u].y), relu2(qa[u].z, qa[u].w), relu2(qb[u].x, qb[u].y),
                            relu2(qb[u].z, qb[u].w)};
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        float2 a = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 4; ++j) a = __ffma2_rn(sr[i][j], q2[j], a);
        val[u][i] = a.x + a.y;
      }
    }
    // transposing butterfly over the 4 k-slices: after it, this lane holds the full sums of pixel p0 + my_px
    float keep[2][5], fin[5];
    const bool odd1 = kl & 1, odd2 = kl & 2;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const float send = odd1 ? val[i][j] : val[i + 2][j];
        const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
        keep[i][j] = (odd1 ? val[i + 2][j] : val[i][j]) + recv;
      }
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const float send = odd2 ? keep[0][j] : keep[1][j];
      const float recv = __shfl_xor_sync(0xffffffffu, send, 2);
      fin[j] = (odd2 ? keep[1][j] : keep[0][j]) + recv;
    }
    const int p = p0 + my_px;
    if (p < HW) {
      const float inv = 1.0f / (fin[4] + eps);
      st4<TO>(out + (static_cast<long long>(f) * HW + p) * Co + g * 32 + cl * 4,
              make_float4(fin[0] * inv, fin[1] * inv, fin[2] * inv, fin[3] * inv));
    }
  }
}

// ---------------------------------------------------------------- ReLU linear attention on warp-level tensor-core MMAs
// Same math as linear_attn_kernel for bf16 tensors, with both products as mma.sync.m16n8k16 (bf16 x bf16 -> fp32):
//   phase 1  S1[c'][c] = sum_px relu(k[px][c']) * [v | 1][px][c]       M = 32 k-channels, N = 40 (32 v + ones + pad), K = pixels
//   phase 2  o[px][c]  = sum_c' relu(q[px][c']) * S1[c'][c]             M = pixels, N = 40, K = 32; out = o[:32] / (o[32] + eps)
// One CTA per (frame, head), 8 warps, each warp walks 16-pixel blocks (cp.async into a private double-buffered tile whose
// 144-byte rows make every ldmatrix conflict-free; k / v fragments come out of ldmatrix.trans because the pixel index is
// the contraction index in phase 1).  q, k, v are bf16 in memory already and the MMAs accumulate in fp32, so phase 1 is
// exactly the SIMT kernel's arithmetic; S1 (fp32) enters phase 2 as hi + lo bf16 halves (two MMAs, relative error 2^-16).
// The SIMT kernel issues ~130 instructions per pixel and was FMA-issue bound at 0.09 of the HBM roofline; this one
// issues ~5 and is bound by the loads.
namespace lamma {
constexpr int ROW = 72;                 // bf16 per staged pixel row: k 32 | v 32 | one + 7 zeros   (144 B)
constexpr int TILE = 16 * ROW;          // one 16-pixel block
constexpr int SB_ROW = 40;              // bf16 per row of the S1 operand tables: 32 k-channels + 8 pad (80 B)
constexpr int SMEM_BYTES = 8 * 2 * TILE * 2 /*kv / q tiles, reused for the per-warp partials of S1*/ + 2 * 40 * SB_ROW * 2 /*hi, lo*/;
static_assert(8 * 32 * 33 * 4 <= 8 * 2 * TILE * 2, "the partials of S1 must fit in the staging tiles");

__device__ __forceinline__ void cp16(void* dst, const void* src, bool valid) {
  const int n = valid ? 16 : 0;  // src-size 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(dst))),
               "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(static_cast<uint32_t>(__cvta_generic_to_shared(p))));
}
__device__ __forceinline__ void ldsm4t(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(static_cast<uint32_t>(__cvta_generic_to_shared(p))));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t relu_bf16x2(uint32_t x) {
  const __nv_bfloat162 z = __floats2bfloat162_rn(0.f, 0.f);
  __nv_bfloat162 v = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&x), z);
  return *reinterpret_cast<uint32_t*>(&v);
}
}  // namespace lamma

__global__ void __launch_bounds__(256, 3) linear_attn_mma_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ ms,
                                                                 bf16* __restrict__ out, int HW, int heads, float eps) {
  using namespace lamma;
  pdl_grid_sync();
  extern __shared__ __align__(16) uint8_t la_smem[];
  bf16* tiles = reinterpret_cast<bf16*>(la_smem);                                  // [8 warps][2][16][ROW]
  bf16* SBhi = reinterpret_cast<bf16*>(la_smem + 8 * 2 * TILE * 2);  // [40 c][SB_ROW]: k-channel contiguous
  bf16* SBlo = SBhi + 40 * SB_ROW;
  const int g = blockIdx.x % (2 * heads);
  const int f = blockIdx.x / (2 * heads);
  const int scale = g / heads, hg = g % heads;
  const int C3 = heads * 96;
  const bf16* src = (scale ? ms : qkv) + static_cast<long long>(f) * HW * C3 + hg * 96;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gi = lane >> 2, ti = lane & 3;        // MMA fragment coordinates
  const int mi = lane >> 3, mr = lane & 7;        // ldmatrix: matrix index / row this lane addresses
  bf16* my = tiles + warp * 2 * TILE;
  const int nblk = (HW + 15) / 16;

  // ------------------------------------------------------------------ phase 1
  auto stage_kv = [&](int blk, int b) {  // 16 px x (k 64 B | v 64 B): 128 16-byte chunks, 4 per lane
    bf16* t = my + b * TILE;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ch = j * 32 + lane, px = ch >> 3, c16 = ch & 7;
      const int p = blk * 16 + px;
      const bool ok = p < HW;
      cp16(t + px * ROW + c16 * 8, src + static_cast<long long>(ok ? p : 0) * C3 + 32 + c16 * 8, ok);
    }
    if (lane < 16) {  // the column of ones (zero for pixels past the end) + zero pad
      const bool ok = blk * 16 + lane < HW;
      *reinterpret_cast<uint4*>(t + lane * ROW + 64) = make_uint4(ok ? 0x00003f80u : 0u, 0u, 0u, 0u);  // bf16 1.0 = 0x3f80
    }
    cp_commit();
  };
  float acc[2][5][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 5; ++b)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;
  int buf = 0;
  if (warp < nblk) stage_kv(warp, 0);
  for (int blk = warp; blk < nblk; blk += 8) {
    const bool more = blk + 8 < nblk;
    if (more) stage_kv(blk + 8, buf ^ 1);
    if (more) cp_wait<1>(); else cp_wait<0>();
    __syncwarp();
    const bf16* t = my + buf * TILE;
    uint32_t a[2][4], b[6][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {  // A = relu(k)^T: matrices (px 0-7 | 8-15) x (c' 0-7 | 8-15) of this 16-channel slab
      ldsm4t(a[mt], t + ((mi >> 1) * 8 + mr) * ROW + mt * 16 + (mi & 1) * 8);
#pragma unroll
      for (int r = 0; r < 4; ++r) a[mt][r] = relu_bf16x2(a[mt][r]);
    }
#pragma unroll
    for (int np = 0; np < 3; ++np) {  // B = [v | 1]: two 8-column tiles per ldmatrix (the 6th tile is pad, never used)
      uint32_t r4[4];
      const int col = 32 + (np * 2 + (mi >> 1)) * 8;
      ldsm4t(r4, t + ((mi & 1) * 8 + mr) * ROW + (col < ROW ? col : 64));
      b[np * 2][0] = r4[0]; b[np * 2][1] = r4[1]; b[np * 2 + 1][0] = r4[2]; b[np * 2 + 1][1] = r4[3];
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 5; ++nt) mma16816(acc[mt][nt], a[mt], b[nt][0], b[nt][1]);
    __syncwarp();
    buf ^= 1;
  }
  // per-warp partials -> shared memory (the staging tiles are dead now), summed in a fixed order: the result does not
  // depend on the order in which warps finish (shared-memory atomics would make repeated calls differ in the last bit)
  __syncthreads();
  float* red = reinterpret_cast<float*>(la_smem);  // [8 warps][32 c'][33 c]
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 5; ++nt) {
      if (nt == 4 && ti != 0) continue;  // tile 4 carries only column 32 (the row sums of relu(k))
      float* r0 = red + (warp * 32 + mt * 16 + gi) * 33 + nt * 8 + 2 * ti;
      r0[0] = acc[mt][nt][0];
      r0[8 * 33] = acc[mt][nt][2];
      if (nt < 4) {
        r0[1] = acc[mt][nt][1];
        r0[8 * 33 + 1] = acc[mt][nt][3];
      }
    }
  __syncthreads();
  for (int i = tid; i < 40 * 32; i += 256) {  // operand tables of phase 2: [c][c'] so that the k index is contiguous
    const int c = i >> 5, kc = i & 31;
    float v = 0.f;
    if (c < 33) {
#pragma unroll
      for (int w = 0; w < 8; ++w) v += red[(w * 32 + kc) * 33 + c];
    }
    const bf16 hi = __float2bfloat16_rn(v);
    SBhi[c * SB_ROW + kc] = hi;
    SBlo[c * SB_ROW + kc] = __float2bfloat16_rn(v - __bfloat162float(hi));
  }
  __syncthreads();  // tables complete; the partials' memory becomes the q staging tiles again
  uint32_t bh[5][4], bl[5][4];  // per 8-column tile: {b0, b1} of k-step 0, {b0, b1} of k-step 1
#pragma unroll
  for (int nt = 0; nt < 5; ++nt) {
    ldsm4(bh[nt], SBhi + (nt * 8 + mr) * SB_ROW + mi * 8);
    ldsm4(bl[nt], SBlo + (nt * 8 + mr) * SB_ROW + mi * 8);
  }

  // ------------------------------------------------------------------ phase 2
  auto stage_q = [&](int blk, int b) {  // 16 px x 64 B: 64 chunks, 2 per lane; rows reuse the 144-byte pitch
    bf16* t = my + b * TILE;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int ch = j * 32 + lane, px = ch >> 2, c16 = ch & 3;
      const int p = blk * 16 + px;
      const bool ok = p < HW;
      cp16(t + px * ROW + c16 * 8, src + static_cast<long long>(ok ? p : 0) * C3 + c16 * 8, ok);
    }
    cp_commit();
  };
  const int Co = 2 * heads * 32;
  buf = 0;
  if (warp < nblk) stage_q(warp, 0);
  for (int blk = warp; blk < nblk; blk += 8) {
    const bool more = blk + 8 < nblk;
    if (more) stage_q(blk + 8, buf ^ 1);
    if (more) cp_wait<1>(); else cp_wait<0>();
    __syncwarp();
    const bf16* t = my + buf * TILE;
    float o[5][4];
#pragma unroll
    for (int nt = 0; nt < 5; ++nt)
#pragma unroll
      for (int c = 0; c < 4; ++c) o[nt][c] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {  // A = relu(q): matrices (px 0-7 | 8-15) x (c' 0-7 | 8-15) of this 16-channel k-step
      uint32_t a[4];
      ldsm4(a, t + ((mi & 1) * 8 + mr) * ROW + ks * 16 + (mi >> 1) * 8);
#pragma unroll
      for (int r = 0; r < 4; ++r) a[r] = relu_bf16x2(a[r]);
#pragma unroll
      for (int nt = 0; nt < 5; ++nt) {
        mma16816(o[nt], a, bh[nt][ks * 2], bh[nt][ks * 2 + 1]);
        mma16816(o[nt], a, bl[nt][ks * 2], bl[nt][ks * 2 + 1]);
      }
    }
    // denominator = column 32 = tile 4, column 0: held by the ti == 0 lane of each row quad
    const float d0 = __shfl_sync(0xffffffffu, o[4][0], lane & ~3), d8 = __shfl_sync(0xffffffffu, o[4][2], lane & ~3);
    const float i0 = 1.0f / (d0 + eps), i8 = 1.0f / (d8 + eps);
    const int p0 = blk * 16 + gi, p8 = p0 + 8;
    bf16* op = out + static_cast<long long>(f) * HW * Co + g * 32 + 2 * ti;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      if (p0 < HW)
        *reinterpret_cast<__nv_bfloat162*>(op + static_cast<long long>(p0) * Co + nt * 8) = __floats2bfloat162_rn(o[nt][0] * i0, o[nt][1] * i0);
      if (p8 < HW)
        *reinterpret_cast<__nv_bfloat162*>(op + static_cast<long long>(p8) * Co + nt * 8) = __floats2bfloat162_rn(o[nt][2] * i8, o[nt][3] * i8);
    }
    __syncwarp();
    buf ^= 1;
  }
}

// ---------------------------------------------------------------- fused multiscale projection, grouped 1x1 on mma.sync
// Phase A (5x5 depthwise sphere conv, sliding window) as in multiscale_fused_kernel; its fp32 results are split into
// bf16 hi + lo halves and stored pixel-major ([64 px][128 channels], 272-byte rows: conflict-free for the 8-byte stores
// of phase A and for ldmatrix).  Phase B (grouped 32 -> 32 1x1: 55 % of the kernel's FMAs) runs as m16n8k16 MMAs per
// (group, 32-pixel half): x_hi w_hi + x_lo w_hi + x_hi w_lo, i.e. fp32-class products (relative error 2^-16) for a
// sixth of the SIMT form's instructions.
namespace msmma {
constexpr int XROW = 136;  // bf16 per pixel row: 128 channels + 8 pad
constexpr int WROW = 40;   // bf16 per weight row: 32 inputs + 8 pad
constexpr int SMEM_BYTES = 2 * 64 * XROW * 2 + 2 * 4 * 32 * WROW * 2;
}  // namespace msmma

__global__ void __launch_bounds__(256, 3) multiscale_fused_mma_kernel(const bf16* __restrict__ in, const float* __restrict__ w5,
                                                                      const float* __restrict__ wg, bf16* __restrict__ out,
                                                                      int n, int H, int W, int C) {
  using namespace lamma;
  using namespace msmma;
  pdl_grid_sync();
  extern __shared__ __align__(16) uint8_t ms_smem[];
  bf16* xh = reinterpret_cast<bf16*>(ms_smem);  // [64 px][XROW]
  bf16* xl = xh + 64 * XROW;
  bf16* wh = xl + 64 * XROW;                    // [4 groups][32 outputs][WROW]: input index contiguous
  bf16* wl = wh + 4 * 32 * WROW;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ncb = (C + 127) / 128;
  const int cb = (blockIdx.x % ncb) * 128;
  const int fy = blockIdx.x / ncb;
  const int y = fy % H, f = fy / H;
  for (int i = tid; i < 4096; i += 256) {  // grouped weights [o][k] -> bf16 hi / lo
    const int g = i >> 10, o = (i >> 5) & 31, k = i & 31;
    float v = 0.f;
    if (cb + g * 32 < C) v = wg[(static_cast<long long>(cb) + g * 32 + o) * 32 + k];
    const bf16 hi = __float2bfloat16_rn(v);
    wh[(g * 32 + o) * WROW + k] = hi;
    wl[(g * 32 + o) * WROW + k] = __float2bfloat16_rn(v - __bfloat162float(hi));
  }
  const int c = cb + lane * 4;
  const bool c_ok = c < C;
  const bf16* base = in + static_cast<long long>(f) * H * W * C + c;
  const int g = warp & 3, half = warp >> 2;
  const bool g_ok = cb + g * 32 < C;
  const int gi = lane >> 2, ti = lane & 3, mi = lane >> 3, mr = lane & 7;
  for (int seg = 0; seg < W; seg += 64) {
    __syncthreads();  // weights staged / previous segment's phase B done with xh / xl
    // ---- phase A
#pragma unroll 1
    for (int it = 0; it < 2; ++it) {
      const int xloc = (it * 8 + warp) * 4;
      const int x0 = seg + xloc;
      float4 acc[4];
#pragma unroll
      for (int o = 0; o < 4; ++o) acc[o] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c_ok && x0 < W) {
        int con[8];
        col_offsets<8>(con, x0, 2, W, C, false);
#pragma unroll
        for (int ky = 0; ky < 5; ++ky) {
          int sy;
          bool rolled;
          sphere_row(y + ky, 2, H, sy, rolled);
          const bool flip = (y == 0 && ky < 2) || (y == H - 1 && ky >= 3);
          float4 wr[5];
#pragma unroll
          for (int kx = 0; kx < 5; ++kx)
            wr[kx] = __ldg(reinterpret_cast<const float4*>(w5 + (ky * 5 + (flip ? 4 - kx : kx)) * C + c));
          const bf16* rowp = base + static_cast<long long>(sy) * W * C;
          int co[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) co[j] = con[j];
          if (rolled) col_offsets<8>(co, x0, 2, W, C, true);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 v = ld4<bf16>(rowp + co[j]);
#pragma unroll
            for (int kx = 0; kx < 5; ++kx) {
              const int o = j - kx;
              if (o >= 0 && o < 4) fma4(acc[o], wr[kx], v);
            }
          }
        }
      }
      // (pixels past the row end / channels past C: zeros, so that no stale bits ever reach the tensor pipe)
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(acc[o].x, acc[o].y), h1 = __floats2bfloat162_rn(acc[o].z, acc[o].w);
        const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
        const __nv_bfloat162 l0 = __floats2bfloat162_rn(acc[o].x - f0.x, acc[o].y - f0.y);
        const __nv_bfloat162 l1 = __floats2bfloat162_rn(acc[o].z - f1.x, acc[o].w - f1.y);
        uint2 hv, lv;
        hv.x = *reinterpret_cast<const uint32_t*>(&h0); hv.y = *reinterpret_cast<const uint32_t*>(&h1);
        lv.x = *reinterpret_cast<const uint32_t*>(&l0); lv.y = *reinterpret_cast<const uint32_t*>(&l1);
        *reinterpret_cast<uint2*>(xh + (xloc + o) * XROW + lane * 4) = hv;
        *reinterpret_cast<uint2*>(xl + (xloc + o) * XROW + lane * 4) = lv;
      }
    }
    __syncthreads();
    // ---- phase B: this warp = (group g, 32-pixel half): two 16-pixel M tiles x four 8-output N tiles x two k-steps
    if (!g_ok || seg + half * 32 >= W) continue;
    uint32_t bh[4][4], bl[4][4];  // per 8-output tile: {b0, b1} of k-step 0, {b0, b1} of k-step 1 (reloaded per segment:
                                  // keeping them across phase A spills)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      ldsm4(bh[nt], wh + (g * 32 + nt * 8 + mr) * WROW + mi * 8);
      ldsm4(bl[nt], wl + (g * 32 + nt * 8 + mr) * WROW + mi * 8);
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const int prow = half * 32 + mt * 16;
      float d[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int q = 0; q < 4; ++q) d[nt][q] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        uint32_t ah[4], al[4];
        const int off = (prow + (mi & 1) * 8 + mr) * XROW + g * 32 + ks * 16 + (mi >> 1) * 8;
        ldsm4(ah, xh + off);
        ldsm4(al, xl + off);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          mma16816(d[nt], ah, bh[nt][ks * 2], bh[nt][ks * 2 + 1]);
          mma16816(d[nt], al, bh[nt][ks * 2], bh[nt][ks * 2 + 1]);
          mma16816(d[nt], ah, bl[nt][ks * 2], bl[nt][ks * 2 + 1]);
        }
      }
      const int p0 = seg + prow + gi, p8 = p0 + 8;
      bf16* op = out + (static_cast<long long>(f) * H + y) * W * C + cb + g * 32 + 2 * ti;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        if (p0 < W) *reinterpret_cast<__nv_bfloat162*>(op + static_cast<long long>(p0) * C + nt * 8) = __floats2bfloat162_rn(d[nt][0], d[nt][1]);
        if (p8 < W) *reinterpret_cast<__nv_bfloat162*>(op + static_cast<long long>(p8) * C + nt * 8) = __floats2bfloat162_rn(d[nt][2], d[nt][3]);
      }
    }
  }
}

// ---------------------------------------------------------------- channel RMSNorm (+residual) on [P, C] rows
// y = x_in * rsqrt(mean(x_in^2) + eps) * w + b ; if resid: resid += y (in place) and the result is also written as T.
template <typename TY, typename T>
__global__ void __launch_bounds__(256) rmsnorm_rows_kernel(const TY* __restrict__ y, int ldy, const float* __restrict__ w,
                                                           const float* __restrict__ b, float eps, float* __restrict__ resid,
                                                           float* __restrict__ out_f32, T* __restrict__ out_t, long long P,
                                                           int C, int relu, int pH, int pW, int pCp) {
  pdl_grid_sync();
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= P) return;
  const TY* yr = y + row * ldy;
  const int c4 = C / 4;
  float ss = 0.f;
  for (int i = lane; i < c4; i += 32) {
    const float4 v = ld4<TY>(yr + i * 4);
    ss += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float r = rsqrtf(ss / C + eps);
  // output row: plain [P, C] rows, or (pCp > 0) the interior of a sphere-padded [n, pH+2, pW+2, pCp] conv input
  // (row -> (frame, y, x) once per row: the 64-bit divisions used to sit inside the channel loop)
  long long obase = row * C;
  if (pCp > 0) {
    const long long fy = row / pW;
    const int xx = static_cast<int>(row - fy * pW);
    const long long ff = fy / pH;
    const int yy = static_cast<int>(fy - ff * pH);
    obase = ((ff * (pH + 2) + yy + 1) * (pW + 2) + xx + 1) * pCp;
  }
  float* rrow = resid != nullptr ? resid + row * C : nullptr;
  float* frow = out_f32 != nullptr ? out_f32 + row * C : nullptr;
  T* trow = out_t != nullptr ? out_t + obase : nullptr;
  for (int i = lane; i < c4; i += 32) {
    const int c = i * 4;
    const float4 v = ld4<TY>(yr + c);
    const float4 ww = __ldg(reinterpret_cast<const float4*>(w + c));
    const float4 bb = __ldg(reinterpret_cast<const float4*>(b + c));
    float4 o = make_float4(v.x * r * ww.x + bb.x, v.y * r * ww.y + bb.y, v.z * r * ww.z + bb.z, v.w * r * ww.w + bb.w);
    if (rrow != nullptr) {
      const float4 h = *reinterpret_cast<const float4*>(rrow + c);
      o.x += h.x; o.y += h.y; o.z += h.z; o.w += h.w;
      *reinterpret_cast<float4*>(rrow + c) = o;
    }
    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    if (frow != nullptr) *reinterpret_cast<float4*>(frow + c) = o;
    if (trow != nullptr) st4<T>(trow + c, o);
  }
}

// ---------------------------------------------------------------- pixel_shuffle(2) of conv output + shortcut
// out[f, 2y+i, 2x+j, c] = conv[f, y, x, 4c+2i+j] + x_in[f, y, x, (4c+2i+j) / rep]     (DCAE.py:519-536)
template <typename TC, typename T>
__global__ void __launch_bounds__(256) pixel_shuffle_kernel(const TC* __restrict__ conv, const float* __restrict__ xin,
                                                            float* __restrict__ out, T* __restrict__ out_t, int n, int H,
                                                            int W, int Cin, int Cout, int rep, int pCp) {
  pdl_grid_sync();
  // one thread = (input pixel, 4 consecutive output channels): 4 float4 of the conv output -> 4 float4 stores, one
  // per sub-pixel (i, j)
  const int c4n = Cout / 4;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(n) * H * W * c4n;
  if (i >= total) return;
  const int c = static_cast<int>(i % c4n) * 4;
  const long long pin = i / c4n;
  const int x = static_cast<int>(pin % W);
  const long long fy = pin / W;
  const int y = static_cast<int>(fy % H);
  const long long f = fy / H;
  const TC* cp = conv + pin * (4LL * Cout) + 4 * c;
  const float* xp = xin + pin * Cin;
  float v[4][4];  // [channel k][sub-pixel 2i+j]
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float4 t = ld4<TC>(cp + 4 * k);
    const int ch = 4 * (c + k);
    v[k][0] = t.x + __ldg(xp + (ch + 0) / rep);
    v[k][1] = t.y + __ldg(xp + (ch + 1) / rep);
    v[k][2] = t.z + __ldg(xp + (ch + 2) / rep);
    v[k][3] = t.w + __ldg(xp + (ch + 3) / rep);
  }
#pragma unroll
  for (int sp = 0; sp < 4; ++sp) {
    const int Y = 2 * y + (sp >> 1), X = 2 * x + (sp & 1);
    const long long opix = (f * (2 * H) + Y) * (2 * W) + X;
    *reinterpret_cast<float4*>(out + opix * Cout + c) = make_float4(v[0][sp], v[1][sp], v[2][sp], v[3][sp]);
    if (out_t != nullptr) {
      // plain rows, or (pCp > 0) the interior of the sphere-padded [n, 2H+2, 2W+2, pCp] input of the next 3x3 conv
      const long long o = pCp > 0 ? ((f * (2 * H + 2) + Y + 1) * (2 * W + 2) + X + 1) * pCp + c : opix * Cout + c;
      st4<T>(out_t + o, make_float4(v[0][sp], v[1][sp], v[2][sp], v[3][sp]));
    }
  }
}

// ---------------------------------------------------------------- conv_in shortcut: x[p, c] += z[f, c / rep, y, x]
template <typename T>
__global__ void in_shortcut_kernel(float* __restrict__ x, T* __restrict__ x_t, PlaneSrc z, int n, int HW, int C, int Cz,
                                   int rep) {
  pdl_grid_sync();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(n) * HW * C;
  if (i >= total) return;
  const int c = static_cast<int>(i % C);
  const long long p = i / C;
  const int f = static_cast<int>(p / HW), pix = static_cast<int>(p % HW);
  const float v = x[i] + plane_load(z, f, c / rep, pix);
  x[i] = v;
  if (x_t != nullptr) x_t[i] = from_f32<T>(v);
}

// ---------------------------------------------------------------- pixel_unshuffle(2) of conv output + shortcut
// DCDownBlock2d (DCAE.py:476-490), H x W = the FINE resolution, Cq = Cout / 4 conv channels, g = 4 Cin / Cout:
//   out[f, y, x, k] = conv[f, 2y + (k&3)/2, 2x + (k&1), k / 4] + mean_{j<g} xu[k g + j],
//   xu[u] = x_in[f, 2y + (u&3)/2, 2x + (u&1), u / 4]                (pixel_unshuffle, then unflatten(1, (-1, g)).mean(2))
template <typename T>
__global__ void __launch_bounds__(256) pixel_unshuffle_kernel(const float* __restrict__ conv, const float* __restrict__ xin,
                                                              float* __restrict__ out, T* __restrict__ out_t, int n, int H,
                                                              int W, int Cin, int Cout, int g, int pCp) {
  pdl_grid_sync();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int Ho = H / 2, Wo = W / 2, Cq = Cout / 4;
  const long long total = static_cast<long long>(n) * Ho * Wo * Cout;
  if (i >= total) return;
  const int k = static_cast<int>(i % Cout);
  const long long opix = i / Cout;
  const int x = static_cast<int>(opix % Wo);
  const long long fy = opix / Wo;
  const int y = static_cast<int>(fy % Ho);
  const long long f = fy / Ho;
  auto fine = [&](int sub) { return (f * H + 2 * y + (sub >> 1)) * W + 2 * x + (sub & 1); };
  float v = conv[fine(k & 3) * Cq + (k >> 2)];
  float sc = 0.f;
  for (int j = 0; j < g; ++j) {
    const int u = k * g + j;
    sc += xin[fine(u & 3) * Cin + (u >> 2)];
  }
  v += sc / static_cast<float>(g);
  out[i] = v;
  if (out_t != nullptr) {
    // plain rows, or (pCp > 0) the interior of the sphere-padded [n, Ho+2, Wo+2, pCp] input of the next 3x3 conv
    const long long o = pCp > 0 ? ((f * (Ho + 2) + y + 1) * (Wo + 2) + x + 1) * pCp + k : i;
    out_t[o] = from_f32<T>(v);
  }
}

// ---------------------------------------------------------------- encoder output shortcut (+ latent normalisation)
// out[f, l, p] (NCHW, conv_out already stored there) += mean_{j<g} x[f, p, l g + j]   (Encoder.forward, DCAE.py:624-627)
// then optionally (v - mean[l]) / std[l] * target  (normalize_transform_3D, dataloader/utils.py:223-231)
__global__ void __launch_bounds__(256) enc_out_shortcut_kernel(float* __restrict__ out, const float* __restrict__ x, int n,
                                                               int HW, int C, int L, int g, const float* __restrict__ mean,
                                                               const float* __restrict__ stdv, float target) {
  pdl_grid_sync();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(n) * L * HW;
  if (i >= total) return;
  const int p = static_cast<int>(i % HW);
  const long long fl = i / HW;
  const int l = static_cast<int>(fl % L);
  const long long f = fl / L;
  const float* xp = x + (f * HW + p) * C + l * g;
  float sc = 0.f;
  for (int j = 0; j < g; ++j) sc += xp[j];
  float v = out[i] + sc / static_cast<float>(g);
  if (mean != nullptr) v = (v - mean[l]) / stdv[l] * target;
  out[i] = v;
}

inline unsigned blocks(long long n, int t = 256) { return static_cast<unsigned>((n + t - 1) / t); }

}  // namespace

template <typename T>
int pad_from_nchw(const PlaneSrc& z, T* out, int n, int C, int H, int W, int Cp, cudaStream_t s) {
  const long long total = static_cast<long long>(n) * (H + 2) * (W + 2) * Cp;
  ProfScope ps(PROF_DEC_PAD, 0.0, static_cast<double>(n) * H * W * C * 4.0 + static_cast<double>(total) * sizeof(T), s);
  LC_CHECK_CUDA(launch_kernel(pad_from_nchw_kernel<T>, blocks(total), 256, 0, s, z, out, n, C, H, W, Cp));
  LC_LAUNCH_CHECK();
  return 0;
}
template <typename T>
int pad_from_nhwc(const float* x, T* out, int n, int C, int H, int W, int Cp, cudaStream_t s) {
  LC_REQUIRE(C % 4 == 0 && Cp % 4 == 0, "pad: channels must be multiples of 4");
  const long long total = static_cast<long long>(n) * (H + 2) * (W + 2) * (Cp / 4);
  ProfScope ps(PROF_DEC_PAD, 0.0, static_cast<double>(n) * H * W * C * 4.0 + static_cast<double>(total) * 4 * sizeof(T), s);
  LC_CHECK_CUDA(launch_kernel(pad_from_nhwc_kernel<T>, blocks(total), 256, 0, s, x, out, n, C, H, W, Cp));
  LC_LAUNCH_CHECK();
  return 0;
}
template <typename T>
int halo_fill(T* buf, int n, int H, int W, int Cp, cudaStream_t s) {
  LC_REQUIRE(Cp % 8 == 0, "halo_fill: Cp must be a multiple of 8");
  const long long total = static_cast<long long>(n) * (2 * (W + 2) + 2 * H) * (Cp / 8);
  ProfScope ps(PROF_DEC_PAD, 0.0, static_cast<double>(total) * 16.0 * sizeof(T), s);
  LC_CHECK_CUDA(launch_kernel(halo_fill_kernel<T>, blocks(total), 256, 0, s, buf, n, H, W, Cp));
  LC_LAUNCH_CHECK();
  return 0;
}
template <typename T>
int dwconv3_glu(const T* in, const float* w, const float* bias, T* out, int n, int H, int W, int C, cudaStream_t s) {
  LC_REQUIRE(C % 8 == 0, "dwconv3_glu: channels must be a multiple of 8");
  constexpr int XT = 4;
  dim3 grid((C / 8 + 31) / 32, (W + 8 * XT - 1) / (8 * XT), n * H);
  // algorithmic: read [P, C] once, write [P, C/2]
  ProfScope ps(PROF_DEC_DWGLU, 0.0, static_cast<double>(n) * H * W * C * 1.5 * sizeof(T), s);
  LC_CHECK_CUDA(launch_kernel(dwconv3_glu_kernel<T, XT>, grid, 256, 0, s, in, w, bias, out, n, H, W, C));
  LC_LAUNCH_CHECK();
  return 0;
}
template <typename T>
int multiscale_launch(const T* in, const float* w5, const float* wg, T* out, int n, int H, int W, int C, unsigned nblk,
                      cudaStream_t s) {
  LC_CHECK_CUDA(launch_kernel(multiscale_fused_kernel<T>, nblk, 256, 0, s, in, w5, wg, out, n, H, W, C));
  return 0;
}
template <>
int multiscale_launch<bf16>(const bf16* in, const float* w5, const float* wg, bf16* out, int n, int H, int W, int C,
                            unsigned nblk, cudaStream_t s) {
  // grouped 1x1 on warp-level MMAs: 10.8 -> 9.8 ms per 80-frame decode, same parity (the depthwise 5x5 of phase A is the
  // bulk of the kernel); LADCAST_B200_MULTISCALE=simt: the all-SIMT kernel
  static const bool mma = [] { const char* e = getenv("LADCAST_B200_MULTISCALE"); return !(e != nullptr && e[0] == 's'); }();
  if (!mma || C % 32 != 0) {
    LC_CHECK_CUDA(launch_kernel(multiscale_fused_kernel<bf16>, nblk, 256, 0, s, in, w5, wg, out, n, H, W, C));
    return 0;
  }
  static PerDevice<bool> attr_set;
  if (!attr_set.here()) {
    LC_CHECK_CUDA(cudaFuncSetAttribute(multiscale_fused_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, msmma::SMEM_BYTES));
    attr_set.here() = true;
  }
  LC_CHECK_CUDA(launch_kernel(multiscale_fused_mma_kernel, nblk, 256, msmma::SMEM_BYTES, s, in, w5, wg, out, n, H, W, C));
  return 0;
}
template <typename T>
int multiscale_fused(const T* in, const float* w5, const float* wg, T* out, int n, int H, int W, int C, cudaStream_t s) {
  LC_REQUIRE(C % 32 == 0, "multiscale projection: channels must be a multiple of 32");
  const long long nblk = static_cast<long long>(n) * H * ((C + 127) / 128);
  LC_REQUIRE(nblk < (1ll << 31), "multiscale projection: too many image rows per call");
  // algorithmic: read qkv [P, C] once, write the multiscale branch [P, C]
  ProfScope ps(PROF_DEC_MS, 2.0 * n * H * W * C * (25.0 + 32.0), static_cast<double>(n) * H * W * C * 2.0 * sizeof(T), s);
  LC_TRY(multiscale_launch<T>(in, w5, wg, out, n, H, W, C, static_cast<unsigned>(nblk), s));
  LC_LAUNCH_CHECK();
  return 0;
}
template <typename T>
int linear_attn_launch(const T* qkv, const T* ms, T* out, int n, int HW, int heads, float eps, cudaStream_t s) {
  LC_CHECK_CUDA(launch_kernel(linear_attn_kernel<T, T>, n * 2 * heads, 256, 0, s, qkv, ms, out, HW, heads, eps));
  return 0;
}
template <>
int linear_attn_launch<bf16>(const bf16* qkv, const bf16* ms, bf16* out, int n, int HW, int heads, float eps, cudaStream_t s) {
  // LADCAST_B200_LINATTN=simt: the register-tiled SIMT kernel (also what the FP32 validation mode runs)
  static const bool simt = [] { const char* e = getenv("LADCAST_B200_LINATTN"); return e != nullptr && e[0] == 's'; }();
  if (simt) {
    LC_CHECK_CUDA(launch_kernel(linear_attn_kernel<bf16, bf16>, n * 2 * heads, 256, 0, s, qkv, ms, out, HW, heads, eps));
    return 0;
  }
  static PerDevice<bool> attr_set;
  if (!attr_set.here()) {
    LC_CHECK_CUDA(cudaFuncSetAttribute(linear_attn_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, lamma::SMEM_BYTES));
    attr_set.here() = true;
  }
  LC_CHECK_CUDA(launch_kernel(linear_attn_mma_kernel, n * 2 * heads, 256, lamma::SMEM_BYTES, s, qkv, ms, out, HW, heads, eps));
  return 0;
}
template <typename T>
int linear_attention(const T* qkv, const T* ms, T* out, int n, int HW, int heads, float eps, cudaStream_t s) {
  // algorithmic: k, v read once (phase 1), q read once (phase 2) of both scales, output [P, 2*heads*32] written
  ProfScope ps(PROF_DEC_LINATTN, 0.0, static_cast<double>(n) * HW * heads * (2 * 96 + 2 * 32) * sizeof(T), s);
  LC_TRY(linear_attn_launch<T>(qkv, ms, out, n, HW, heads, eps, s));
  LC_LAUNCH_CHECK();
  return 0;
}
template <typename TY, typename T>
int rmsnorm_rows(const TY* y, int ldy, const float* w, const float* b, float eps, float* resid, float* out_f32, T* out_t,
                 long long P, int C, int relu, cudaStream_t s, int pH, int pW, int pCp) {
  LC_REQUIRE(C % 4 == 0 && ldy % 4 == 0, "rmsnorm: C and the row pitch must be multiples of 4");
  // algorithmic: read y; read + write the residual; write the f32 / T copies that were asked for
  ProfScope ps(PROF_DEC_NORM, 0.0,
               static_cast<double>(P) * C * (sizeof(TY) + (resid ? 8.0 : 0.0) + (out_f32 ? 4.0 : 0.0) + (out_t ? sizeof(T) : 0.0)), s);
  LC_CHECK_CUDA(launch_kernel(rmsnorm_rows_kernel<TY, T>, blocks(P, 8), 256, 0, s, y, ldy, w, b, eps, resid, out_f32, out_t, P, C, relu, pH, pW, pCp));
  LC_LAUNCH_CHECK();
  return 0;
}
template <typename TC, typename T>
int pixel_shuffle_shortcut(const TC* conv, const float* xin, float* out, T* out_t, int n, int H, int W, int Cin,
                           int Cout, cudaStream_t s, int pCp) {
  LC_REQUIRE(Cout % 4 == 0, "pixel_shuffle: C_out must be a multiple of 4");
  const long long total = static_cast<long long>(n) * H * W * (Cout / 4);
  ProfScope ps(PROF_DEC_SHUFFLE, 0.0,
               static_cast<double>(n) * H * W * (4.0 * Cout * (sizeof(TC) + 4.0 + (out_t ? sizeof(T) : 0.0)) + Cin * 4.0), s);
  LC_CHECK_CUDA(launch_kernel(pixel_shuffle_kernel<TC, T>, blocks(total), 256, 0, s, conv, xin, out, out_t, n, H, W, Cin, Cout, 4 * Cout / Cin, pCp));
  LC_LAUNCH_CHECK();
  return 0;
}
template <typename T>
int pixel_unshuffle_shortcut(const float* conv, const float* xin, float* out, T* out_t, int n, int H, int W, int Cin,
                             int Cout, cudaStream_t s, int pCp) {
  LC_REQUIRE(Cout % 4 == 0 && (4 * Cin) % Cout == 0 && H % 2 == 0 && W % 2 == 0, "pixel_unshuffle: bad geometry");
  const long long total = static_cast<long long>(n) * (H / 2) * (W / 2) * Cout;
  ProfScope ps(PROF_DEC_SHUFFLE, 0.0,
               static_cast<double>(n) * H * W * (Cout / 4 + Cin) * 4.0 + static_cast<double>(total) * (4.0 + (out_t ? sizeof(T) : 0.0)), s);
  LC_CHECK_CUDA(launch_kernel(pixel_unshuffle_kernel<T>, blocks(total), 256, 0, s, conv, xin, out, out_t, n, H, W, Cin, Cout, 4 * Cin / Cout, pCp));
  LC_LAUNCH_CHECK();
  return 0;
}
int enc_out_shortcut(float* out, const float* x, int n, int HW, int C, int L, const float* mean, const float* stdv,
                     float target, cudaStream_t s) {
  LC_REQUIRE(C % L == 0, "encoder out shortcut needs C divisible by latent_channels");
  const long long total = static_cast<long long>(n) * L * HW;
  ProfScope ps(PROF_DEC_PAD, 0.0, static_cast<double>(total) * 8.0 + static_cast<double>(n) * HW * C * 4.0, s);
  LC_CHECK_CUDA(launch_kernel(enc_out_shortcut_kernel, blocks(total), 256, 0, s, out, x, n, HW, C, L, C / L, mean, stdv, target));
  LC_LAUNCH_CHECK();
  return 0;
}
template <typename T>
int in_shortcut(float* x, T* x_t, const PlaneSrc& z, int n, int HW, int C, int Cz, cudaStream_t s) {
  const long long total = static_cast<long long>(n) * HW * C;
  ProfScope ps(PROF_DEC_PAD, 0.0, static_cast<double>(total) * (8.0 + sizeof(T)) + static_cast<double>(n) * HW * Cz * 4.0, s);
  LC_CHECK_CUDA(launch_kernel(in_shortcut_kernel<T>, blocks(total), 256, 0, s, x, x_t, z, n, HW, C, Cz, C / Cz));
  LC_LAUNCH_CHECK();
  return 0;
}

#define LC_INST(T)                                                                                                   \
  template int pad_from_nchw<T>(const PlaneSrc&, T*, int, int, int, int, int, cudaStream_t);                          \
  template int pad_from_nhwc<T>(const float*, T*, int, int, int, int, int, cudaStream_t);                            \
  template int halo_fill<T>(T*, int, int, int, int, cudaStream_t);                                                   \
  template int dwconv3_glu<T>(const T*, const float*, const float*, T*, int, int, int, int, cudaStream_t);           \
  template int multiscale_fused<T>(const T*, const float*, const float*, T*, int, int, int, int, cudaStream_t);      \
  template int linear_attention<T>(const T*, const T*, T*, int, int, int, float, cudaStream_t);                      \
  template int rmsnorm_rows<T, T>(const T*, int, const float*, const float*, float, float*, float*, T*, long long,   \
                                  int, int, cudaStream_t, int, int, int);                                            \
  template int pixel_shuffle_shortcut<T, T>(const T*, const float*, float*, T*, int, int, int, int, int,             \
                                            cudaStream_t, int);                                                      \
  template int pixel_unshuffle_shortcut<T>(const float*, const float*, float*, T*, int, int, int, int, int,          \
                                           cudaStream_t, int);                                                       \
  template int in_shortcut<T>(float*, T*, const PlaneSrc&, int, int, int, int, cudaStream_t);
LC_INST(float)
LC_INST(bf16)
#undef LC_INST
// fp32 input rows with a bf16 target (norm_out reads the fp32 residual stream in the bf16 mode)
template int rmsnorm_rows<float, bf16>(const float*, int, const float*, const float*, float, float*, float*, bf16*, long long,
                                       int, int, cudaStream_t, int, int, int);

}  // namespace lc
