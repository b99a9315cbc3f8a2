// HBM-bound kernels of the denoiser: LayerNorm + modulation, QK RMSNorm + RoPE, patchify, timestep embedding,
// token pooling, gated residual, temb combine; and the fused scheduler steps.  All vectorised (128-bit) with
// warp-shuffle reductions; fp32 statistics regardless of the storage type.
#include <cuda_fp16.h>
#include <cstdlib>
#include "kernels.h"
#include "ptx.cuh"

namespace lc {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T>
__device__ __forceinline__ void store4(T* p, float a, float b, float c, float d);
template <>
__device__ __forceinline__ void store4<float>(float* p, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
template <>
__device__ __forceinline__ void store4<bf16>(bf16* p, float a, float b, float c, float d) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&lo);
  u.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(p) = u;
}
template <typename T>
__device__ __forceinline__ float4 load4(const T* p);
template <>
__device__ __forceinline__ float4 load4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <>
__device__ __forceinline__ float4 load4<bf16>(const bf16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
  const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
  const float2 a = __bfloat1622float2(lo), b = __bfloat1622float2(hi);
  return make_float4(a.x, a.y, b.x, b.y);
}

// ---------------------------------------------------------------- LayerNorm (+ AdaLN modulation | affine)
// One warp per row at a time, the row (d = NV*128 floats) lives in registers; two-pass variance (biased), as
// nn.LayerNorm.  Each warp walks over RPW consecutive rows with the next row's loads in flight while the current row is
// normalised (in place) and stored (the one-row-per-warp version was DRAM-latency bound: 35 % of HBM peak).
// Round 2 tried two shared-memory rings in front of this (cp.async per lane; one TMA bulk copy per row with a per-warp
// mbarrier ring): both were SLOWER (3.0-3.4 TB/s stand-alone against 4.1 TB/s here; tools/bench_ln.py) — plain 128-bit
// loads from 16 warps per SM keep more bytes in flight than 8 warps x 3 rows — so the register pipeline stays, now
// without spills (the row is normalised in its buffer; d = 2048 runs one CTA per SM with the full register file).
// Reference: AdaLayerNormZero/ZeroSingle/Continuous (diffusers), LaDCast_3D_model.py:287-302, 524-552, 1044.
template <typename T, int NV>
__device__ __forceinline__ void ln_row(float4 (&v)[NV], int row, int lane, T* __restrict__ out, int d, float eps,
                                       int rows_per_sample, int seg_rows, int seg_rows_per_sample,
                                       const float* __restrict__ scale, const float* __restrict__ shift,
                                       long long mod_stride, const float* __restrict__ w, const float* __restrict__ b) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) / d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
    q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
  }
  const float rstd = rsqrtf(warp_sum(q) / d + eps);
  // rows >= seg_rows (> 0) are a second stream with seg_rows_per_sample rows per sample (pred | cond in one launch)
  const int sample = (seg_rows > 0 && row >= seg_rows) ? (row - seg_rows) / seg_rows_per_sample : row / rows_per_sample;
  T* orow = out + static_cast<long long>(row) * d;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    float4 y = make_float4(v[i].x * rstd, v[i].y * rstd, v[i].z * rstd, v[i].w * rstd);
    if (scale != nullptr) {
      const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + sample * mod_stride + c));
      const float4 sh = __ldg(reinterpret_cast<const float4*>(shift + sample * mod_stride + c));
      y.x = y.x * (1.f + sc.x) + sh.x; y.y = y.y * (1.f + sc.y) + sh.y;
      y.z = y.z * (1.f + sc.z) + sh.z; y.w = y.w * (1.f + sc.w) + sh.w;
    } else if (w != nullptr) {
      const float4 ww = __ldg(reinterpret_cast<const float4*>(w + c));
      const float4 bb = __ldg(reinterpret_cast<const float4*>(b + c));
      y.x = y.x * ww.x + bb.x; y.y = y.y * ww.y + bb.y; y.z = y.z * ww.z + bb.z; y.w = y.w * ww.w + bb.w;
    }
    store4<T>(orow + c, y.x, y.y, y.z, y.w);
  }
}

constexpr int LN_RPW = 4;  // max rows per warp (8 measured slower: too few blocks for the 9000-row streams); the launcher
                           // picks 4, 2 or 1 per launch so that small row counts (2-3 members per GPU) still fill whole waves

// Warps per CTA (WPB): 8, two CTAs per SM, up to d = 1536.  d = 2048 needs ~160 registers per thread for the double-
// buffered row, which allows 384 threads per SM: three 4-warp CTAs (12 warps in flight) instead of one 8-warp CTA.
template <typename T, int NV, int WPB, int MIN_CTAS>
__global__ void __launch_bounds__(WPB * 32, MIN_CTAS) layernorm_kernel(const float* __restrict__ x, T* __restrict__ out, int M, int d,
                                                           float eps, int rows_per_sample, int seg_rows,
                                                           int seg_rows_per_sample, const float* __restrict__ scale,
                                                           const float* __restrict__ shift, long long mod_stride,
                                                           const float* __restrict__ w, const float* __restrict__ b, int rpw) {
  pdl_grid_sync();
  const int row0 = (blockIdx.x * WPB + (threadIdx.x >> 5)) * rpw;
  const int lane = threadIdx.x & 31;
  if (row0 >= M) return;
  const int nrows = min(rpw, M - row0);
  float4 buf[2][NV];
  const float* xr = x + static_cast<long long>(row0) * d + lane * 4;
#pragma unroll
  for (int i = 0; i < NV; ++i) buf[0][i] = *reinterpret_cast<const float4*>(xr + i * 128);
#pragma unroll
  for (int r = 0; r < LN_RPW; ++r) {
    if (r < nrows) {
      if (r + 1 < nrows) {
#pragma unroll
        for (int i = 0; i < NV; ++i)
          buf[(r + 1) & 1][i] = *reinterpret_cast<const float4*>(xr + static_cast<long long>(r + 1) * d + i * 128);
      }
      ln_row<T, NV>(buf[r & 1], row0 + r, lane, out, d, eps, rows_per_sample, seg_rows, seg_rows_per_sample, scale, shift,
                    mod_stride, w, b);
    }
  }
}

// ---------------------------------------------------------------- per-head RMSNorm(q,k) + RoPE, in place
// One warp per (token row, q|k, head); lane owns 4 consecutive features = 2 rotation pairs (head_dim == 128).
// Reference: LaDCast_3D_model.py:103-169 + diffusers RMSNorm / apply_rotary_emb (interleaved pairs).
template <typename T>
__global__ void __launch_bounds__(256) qk_norm_rope_kernel(T* __restrict__ qkv, long long ld, int B, int S, int heads,
                                                           float eps, RopeSeg s0, RopeSeg s1, int nseg) {
  pdl_grid_sync();
  const long long gw = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const long long total = static_cast<long long>(B) * S * 2 * heads;
  if (gw >= total) return;
  const int head = static_cast<int>(gw % heads);
  const int which = static_cast<int>((gw / heads) % 2);  // 0 = q, 1 = k
  const long long row = gw / (2 * heads);
  const int tok = static_cast<int>(row % S);
  const RopeSeg& sg = (nseg > 1 && tok >= s1.start) ? s1 : s0;
  T* p = qkv + row * ld + static_cast<long long>(which) * heads * 128 + head * 128 + lane * 4;
  float4 v = load4<T>(p);
  const float ss = warp_sum((v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w));
  const float r = rsqrtf(ss * (1.0f / 128.0f) + eps);
  const float4 w = __ldg(reinterpret_cast<const float4*>((which ? sg.wk : sg.wq) + lane * 4));
  v.x = v.x * r * w.x; v.y = v.y * r * w.y; v.z = v.z * r * w.z; v.w = v.w * r * w.w;
  if (sg.cos != nullptr) {
    const long long t = static_cast<long long>(tok - sg.start) * 128 + lane * 4;
    const float4 c = __ldg(reinterpret_cast<const float4*>(sg.cos + t));
    const float4 s = __ldg(reinterpret_cast<const float4*>(sg.sin + t));
    const float4 o = make_float4(v.x * c.x - v.y * s.x, v.y * c.y + v.x * s.y, v.z * c.z - v.w * s.z,
                                 v.w * c.w + v.z * s.w);
    v = o;
  }
  store4<T>(p, v.x, v.y, v.z, v.w);
}

// bf16 production variant: 16 lanes per (token, head), 8 features (16 B) per lane -> 128-bit loads/stores.  A thread
// normalises and rotates q AND k of its (token, head, 8-feature slice): both loads are issued up front and the cos /
// sin rows are fetched once for the pair; all index arithmetic is 32-bit.
__device__ __forceinline__ void unpack8(const uint4& raw, float (&v)[8]) {
  const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h2[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  uint4 o;
  __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
  for (int i = 0; i < 4; ++i) o2[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  return o;
}
__global__ void __launch_bounds__(256) qk_norm_rope_bf16_kernel(bf16* __restrict__ qkv, long long ld, int B, int S,
                                                                int heads, float eps, RopeSeg s0, RopeSeg s1, int nseg) {
  pdl_grid_sync();
  const unsigned total = static_cast<unsigned>(B) * S * heads;  // (token row, head) units
  const unsigned unit = blockIdx.x * 16u + (threadIdx.x >> 4);
  const int sub = threadIdx.x & 15;
  const bool live = unit < total;  // no early return: every lane takes part in the shuffles below
  const unsigned u = live ? unit : total - 1;
  const unsigned row = u / heads;
  const int head = static_cast<int>(u - row * heads);
  const int tok = static_cast<int>(row % S);
  const RopeSeg& sg = (nseg > 1 && tok >= s1.start) ? s1 : s0;
  bf16* pq = qkv + static_cast<long long>(row) * ld + head * 128 + sub * 8;
  bf16* pk = pq + heads * 128;
  const uint4 raw_q = *reinterpret_cast<const uint4*>(pq);
  const uint4 raw_k = *reinterpret_cast<const uint4*>(pk);
  float cs[8], sn[8];
  const bool rope = sg.cos != nullptr;
  if (rope && sg.cs != nullptr) {
    // packed table: 4 rotation pairs = one 16-byte load instead of 64 bytes of fp32 cos + sin
    const uint4 c4 = __ldg(reinterpret_cast<const uint4*>(sg.cs + (tok - sg.start) * 64 + sub * 4));
    const uint32_t cw[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&cw[i]));
      cs[2 * i] = cs[2 * i + 1] = f.x;
      sn[2 * i] = sn[2 * i + 1] = f.y;
    }
  } else if (rope) {
    const int t = (tok - sg.start) * 128 + sub * 8;
    const float4 c0 = __ldg(reinterpret_cast<const float4*>(sg.cos + t)), c1 = __ldg(reinterpret_cast<const float4*>(sg.cos + t + 4));
    const float4 n0 = __ldg(reinterpret_cast<const float4*>(sg.sin + t)), n1 = __ldg(reinterpret_cast<const float4*>(sg.sin + t + 4));
    cs[0] = c0.x; cs[1] = c0.y; cs[2] = c0.z; cs[3] = c0.w; cs[4] = c1.x; cs[5] = c1.y; cs[6] = c1.z; cs[7] = c1.w;
    sn[0] = n0.x; sn[1] = n0.y; sn[2] = n0.z; sn[3] = n0.w; sn[4] = n1.x; sn[5] = n1.y; sn[6] = n1.z; sn[7] = n1.w;
  }
  float q[8], k[8];
  unpack8(raw_q, q);
  unpack8(raw_k, k);
  float sq = 0.f, sk = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { sq = fmaf(q[i], q[i], sq); sk = fmaf(k[i], k[i], sk); }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) {
    sq += __shfl_xor_sync(0xffffffffu, sq, o);
    sk += __shfl_xor_sync(0xffffffffu, sk, o);
  }
  const float rq = rsqrtf(sq * (1.0f / 128.0f) + eps), rk = rsqrtf(sk * (1.0f / 128.0f) + eps);
  {
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(sg.wq + sub * 8)), a1 = __ldg(reinterpret_cast<const float4*>(sg.wq + sub * 8 + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(sg.wk + sub * 8)), b1 = __ldg(reinterpret_cast<const float4*>(sg.wk + sub * 8 + 4));
    const float wq[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const float wk[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) { q[i] *= rq * wq[i]; k[i] *= rk * wk[i]; }
  }
  if (rope) {
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      const float a = q[i], b = q[i + 1], c = k[i], d = k[i + 1];
      q[i] = a * cs[i] - b * sn[i];
      q[i + 1] = b * cs[i + 1] + a * sn[i + 1];
      k[i] = c * cs[i] - d * sn[i];
      k[i + 1] = d * cs[i + 1] + c * sn[i + 1];
    }
  }
  if (live) {
    *reinterpret_cast<uint4*>(pq) = pack8(q);
    *reinterpret_cast<uint4*>(pk) = pack8(k);
  }
}

// ---------------------------------------------------------------- patchify: [B,C,THW] f32 -> [B*THW, Kp] T (zero pad)
// Reference: HunyuanVideoPatchEmbed flatten(2).transpose(1,2) with patch (1,1,1): token n = t*HW + h*W + w
// (embeddings.py:56-59).  Pure index permutation -> bit-exact in the fp32 mode.
template <typename T>
__global__ void __launch_bounds__(256) patchify_kernel(const float* __restrict__ x, T* __restrict__ out, int C, int THW,
                                                       int Kp) {
  pdl_grid_sync();
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, n = n0 + tx;
    tile[i][tx] = (c < C && n < THW) ? x[(static_cast<long long>(b) * C + c) * THW + n] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int n = n0 + i, c = c0 + tx;
    if (n < THW && c < Kp) out[(static_cast<long long>(b) * THW + n) * Kp + c] = from_f32<T>(tile[tx][i]);
  }
}

// ---------------------------------------------------------------- Timesteps(256, flip_sin_to_cos, shift 0): [cos | sin]
template <typename T>
__global__ void timestep_embed_kernel(const float* __restrict__ t, int n_t, int B, T* __restrict__ out) {
  pdl_grid_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 128) return;
  const int b = i >> 7, k = i & 127;
  const float f = expf(-logf(10000.0f) * static_cast<float>(k) / 128.0f);
  const float arg = t[n_t == 1 ? 0 : b] * f;
  out[b * 256 + k] = from_f32<T>(cosf(arg));
  out[b * 256 + 128 + k] = from_f32<T>(sinf(arg));
}

// ---------------------------------------------------------------- mean over tokens: [B,N,d] f32 -> [B,d] T
template <typename T>
__global__ void __launch_bounds__(256) token_mean_kernel(const float* __restrict__ x, int N, int d, T* __restrict__ out) {
  pdl_grid_sync();
  __shared__ float part[8][32 * 4 + 4];
  const int b = blockIdx.y;
  const int c = blockIdx.x * 128 + (threadIdx.x & 31) * 4;
  const int ty = threadIdx.x >> 5;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < d) {
    const float* p = x + static_cast<long long>(b) * N * d + c;
    for (int n = ty; n < N; n += 8) {
      const float4 v = *reinterpret_cast<const float4*>(p + static_cast<long long>(n) * d);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  float* pp = &part[ty][(threadIdx.x & 31) * 4];
  pp[0] = acc.x; pp[1] = acc.y; pp[2] = acc.z; pp[3] = acc.w;
  __syncthreads();
  if (ty == 0 && c < d) {
    float r[4] = {0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < 8; ++j)
      for (int q = 0; q < 4; ++q) r[q] += part[j][(threadIdx.x & 31) * 4 + q];
    const float inv = 1.0f / N;
    for (int q = 0; q < 4; ++q) out[static_cast<long long>(b) * d + c + q] = from_f32<T>(r[q] * inv);
  }
}

// ---------------------------------------------------------------- h += a * gate[b]  (refiner attention has no out-proj)
template <typename T>
__global__ void __launch_bounds__(256) gated_add_kernel(float* __restrict__ h, const T* __restrict__ a,
                                                        const float* __restrict__ gate, long long gate_stride,
                                                        long long n4, int d, int rows_per_sample) {
  pdl_grid_sync();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const long long e = i * 4;
  const long long row = e / d;
  const int c = static_cast<int>(e - row * d);
  const int sample = static_cast<int>(row / rows_per_sample);
  float4 hv = *reinterpret_cast<float4*>(h + e);
  const float4 av = load4<T>(a + e);
  const float4 g = __ldg(reinterpret_cast<const float4*>(gate + sample * gate_stride + c));
  hv.x += av.x * g.x; hv.y += av.y * g.y; hv.z += av.z * g.z; hv.w += av.w * g.w;
  *reinterpret_cast<float4*>(h + e) = hv;
}

template <typename T>
__global__ void temb_combine_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                    const float* __restrict__ sc, const float* __restrict__ sh, long long sc_stride,
                                    int B, int d, float* __restrict__ out_f32, T* __restrict__ out_silu) {
  pdl_grid_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * d) return;
  const int r = i / d, c = i - r * d;
  float v = a[i] + (b ? b[i] : 0.f);
  if (sc) v = v * (1.f + sc[r * sc_stride + c]) + sh[r * sc_stride + c];
  if (out_f32) out_f32[i] = v;
  if (out_silu) out_silu[i] = from_f32<T>(v / (1.f + expf(-v)));
}

template <typename T>
__global__ void cast_kernel(const float* __restrict__ x, T* __restrict__ out, long long n) {
  pdl_grid_sync();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = from_f32<T>(x[i]);
}

// ---------------------------------------------------------------- scheduler steps
// DPM-Solver++ (1st order / 2M midpoint) of diffusers' EDMDPMSolverMultistepScheduler.step with
// precondition_outputs and the NEXT step's scale_model_input fused (pipeline_AR.py:87-102).
__global__ void __launch_bounds__(256) dpmpp2m_kernel(const float* __restrict__ f, float* __restrict__ x,
                                                      float* __restrict__ x0_prev, float* __restrict__ x_in_next,
                                                      long long n4, SchedCoef c) {
  pdl_grid_sync();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 fv = reinterpret_cast<const float4*>(f)[i];
  float4 xv = reinterpret_cast<float4*>(x)[i];
  float4 x0 = make_float4(c.c_skip * xv.x + c.c_out * fv.x, c.c_skip * xv.y + c.c_out * fv.y,
                          c.c_skip * xv.z + c.c_out * fv.z, c.c_skip * xv.w + c.c_out * fv.w);
  float4 xn = make_float4(c.a_x * xv.x + c.a_x0 * x0.x, c.a_x * xv.y + c.a_x0 * x0.y, c.a_x * xv.z + c.a_x0 * x0.z,
                          c.a_x * xv.w + c.a_x0 * x0.w);
  if (c.a_d != 0.f) {
    const float4 p = reinterpret_cast<const float4*>(x0_prev)[i];
    xn.x += c.a_d * (x0.x - p.x); xn.y += c.a_d * (x0.y - p.y); xn.z += c.a_d * (x0.z - p.z); xn.w += c.a_d * (x0.w - p.w);
  }
  reinterpret_cast<float4*>(x0_prev)[i] = x0;
  reinterpret_cast<float4*>(x)[i] = xn;
  if (x_in_next != nullptr)
    reinterpret_cast<float4*>(x_in_next)[i] =
        make_float4(xn.x * c.c_in_next, xn.y * c.c_in_next, xn.z * c.c_in_next, xn.w * c.c_in_next);
}

// EDM Heun sampler in float64 (edm_sampler.py:65-113).  phase 0 = Euler predictor, phase 1 = trapezoid corrector.
__global__ void __launch_bounds__(256) heun_kernel(const float* __restrict__ f, double* __restrict__ x,
                                                   double* __restrict__ x_hat, double* __restrict__ d_cur,
                                                   float* __restrict__ x_in_next, long long n, int phase, double t_cur,
                                                   double t_next, double c_skip, double c_out, double c_in_next) {
  pdl_grid_sync();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double fv = static_cast<double>(f[i]);
  double xn;
  if (phase == 0) {
    const double xh = x[i];
    const double den = c_skip * xh + c_out * fv;
    const double dc = (xh - den) / t_cur;
    xn = xh + (t_next - t_cur) * dc;
    x_hat[i] = xh;
    d_cur[i] = dc;
  } else {
    const double xc = x[i];
    const double den = c_skip * xc + c_out * fv;
    const double dp = (xc - den) / t_next;
    xn = x_hat[i] + (t_next - t_cur) * (0.5 * d_cur[i] + 0.5 * dp);
  }
  x[i] = xn;
  if (x_in_next != nullptr) x_in_next[i] = static_cast<float>(xn * c_in_next);
}

// x_in = x * c_in  (scale_model_input of step 0, pipeline_AR.py:90)
__global__ void __launch_bounds__(256) scale_kernel(const float* __restrict__ x, float* __restrict__ out, long long n4,
                                                    float c) {
  pdl_grid_sync();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = reinterpret_cast<const float4*>(x)[i];
  reinterpret_cast<float4*>(out)[i] = make_float4(v.x * c, v.y * c, v.z * c, v.w * c);
}

// Heun prologue (edm_sampler.py:44-46, 56-58): x = float64(noise) * t_0 ; x_in = float32(x * c_in(t_0))
__global__ void __launch_bounds__(256) heun_init_kernel(const float* __restrict__ noise, double* __restrict__ x,
                                                        float* __restrict__ x_in, long long n, double t0, double c_in) {
  pdl_grid_sync();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double v = static_cast<double>(noise[i]) * t0;
  x[i] = v;
  x_in[i] = static_cast<float>(v * c_in);
}

// Stochastic churn of the Heun sampler (edm_sampler.py:67-76, deterministic=False): x_hat = x_cur + k * noise with
// k = sqrt(t_hat^2 - t_cur^2) * S_noise and float64 noise (randn_like of the fp64 state); x_in = float32(x_hat * c_in(t_hat)).
__global__ void __launch_bounds__(256) heun_churn_kernel(double* __restrict__ x, const double* __restrict__ noise,
                                                         float* __restrict__ x_in, long long n, double k, double c_in) {
  pdl_grid_sync();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double v = x[i] + k * noise[i];
  x[i] = v;
  x_in[i] = static_cast<float>(v * c_in);
}

// AR feedback of roll_out_serial (pipelines/utils.py:560-585): from one sampler output [B, C, T, hw] (normalised
// latents) write (a) the next step's conditioning = the last t_in frames [B, C, t_in, hw] (unchanged values) and
// (b) optionally the de-normalised latents (x / target_std) * std[c] + mean[c] in the same layout
// (inverse_normalize_transform_3D, dataloader/utils.py:233-240; separate div / mul / add like the eager reference).
__global__ void __launch_bounds__(256) latent_feedback_kernel(const float* __restrict__ s, float* __restrict__ known,
                                                              float* __restrict__ phys, const float* __restrict__ mean,
                                                              const float* __restrict__ stdv, float target, int C, int T,
                                                              int t_in, int hw2, long long n2) {
  pdl_grid_sync();
  // one thread = two consecutive pixels of a (b, c, t) plane (h*w = 450 is even but not a multiple of 4)
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n2) return;
  const int p = static_cast<int>(i % hw2);
  long long r = i / hw2;
  const int t = static_cast<int>(r % T);
  r /= T;
  const int c = static_cast<int>(r % C);
  const long long b = r / C;
  const float2 v = reinterpret_cast<const float2*>(s)[i];
  if (known != nullptr && t >= T - t_in)
    reinterpret_cast<float2*>(known)[((b * C + c) * t_in + (t - (T - t_in))) * hw2 + p] = v;
  if (phys != nullptr) {
    const float sd = __ldg(stdv + c), mu = __ldg(mean + c);
    reinterpret_cast<float2*>(phys)[i] = make_float2(__fadd_rn(__fmul_rn(__fdiv_rn(v.x, target), sd), mu),
                                                     __fadd_rn(__fmul_rn(__fdiv_rn(v.y, target), sd), mu));
  }
}

}  // namespace

template <typename T>
int layernorm_modulate(const float* x, T* out, int M, int d, float eps, int rows_per_sample, const float* scale,
                       const float* shift, long long mod_stride, const float* w, const float* b, cudaStream_t s,
                       int seg_rows, int seg_rows_per_sample) {
  LC_REQUIRE(d % 128 == 0 && d <= 2048, "layernorm: d must be a multiple of 128, <= 2048");
  const int nv = d / 128;
  // rows per warp: the largest of 4 / 2 / 1 whose grid fills >= 90 % of its last wave (else the best-filling one)
  // LADCAST_B200_LN_WPB=8: the round-1 form for d > 1536 (one 8-warp CTA per SM)
  static const bool wide8 = [] { const char* e = getenv("LADCAST_B200_LN_WPB"); return e != nullptr && e[0] == '8'; }();
  const bool small_cta = nv > 12 && !wide8;
  const int wpb = small_cta ? 4 : 8;
  const int slots = num_sms() * (nv <= 12 ? 2 : small_cta ? 3 : 1);
  int rpw = LN_RPW;
  double best = -1.0;
  for (int cand = LN_RPW; cand >= 1; cand >>= 1) {
    const int blocks = ceil_div(M, wpb * cand);
    const double eff = static_cast<double>(blocks) / (static_cast<double>(ceil_div(blocks, slots)) * slots);
    if (eff >= 0.9) { rpw = cand; break; }
    if (eff > best) { best = eff; rpw = cand; }
  }
  dim3 grid(ceil_div(M, wpb * rpw));
  ProfScope ps(PROF_LN, 0.0, static_cast<double>(M) * d * (4 + sizeof(T)), s);
#define LC_LN_LAUNCH(NV, WPB, MINC)                                                                                    \
  LC_CHECK_CUDA(launch_kernel(layernorm_kernel<T, NV, WPB, MINC>, grid, WPB * 32, 0, s, x, out, M, d, eps, rows_per_sample, \
                              seg_rows, seg_rows_per_sample, scale, shift, mod_stride, w, b, rpw))
#define LC_LN_CASE(NV)                   \
  case NV:                               \
    LC_LN_LAUNCH(NV, 8, 2);              \
    break;
#define LC_LN_CASE_BIG(NV)                                  \
  case NV:                                                  \
    if (small_cta) { LC_LN_LAUNCH(NV, 4, 3); }              \
    else { LC_LN_LAUNCH(NV, 8, 1); }                        \
    break;
  switch (nv) {
    LC_LN_CASE(1) LC_LN_CASE(2) LC_LN_CASE(3) LC_LN_CASE(4) LC_LN_CASE(5) LC_LN_CASE(6) LC_LN_CASE(7) LC_LN_CASE(8)
    LC_LN_CASE(9) LC_LN_CASE(10) LC_LN_CASE(11) LC_LN_CASE(12) LC_LN_CASE_BIG(13) LC_LN_CASE_BIG(14) LC_LN_CASE_BIG(15)
    LC_LN_CASE_BIG(16)
  }
#undef LC_LN_CASE_BIG
#undef LC_LN_LAUNCH
#undef LC_LN_CASE
  LC_LAUNCH_CHECK();
  return 0;
}

template <typename T>
int qk_norm_rope(T* qkv, long long ld, int B, int S, int heads, int head_dim, float eps, const RopeSeg* segs, int nseg,
                 cudaStream_t s) {
  LC_REQUIRE(head_dim == 128, "qk_norm_rope: head_dim must be 128");
  LC_REQUIRE(nseg == 1 || nseg == 2, "qk_norm_rope: 1 or 2 segments");
  const long long total = static_cast<long long>(B) * S * 2 * heads;
  RopeSeg s1 = nseg > 1 ? segs[1] : segs[0];
  ProfScope ps(PROF_ROPE, 0.0, static_cast<double>(total) * 128 * sizeof(T) * 2, s);  // q and k read + written once
  if (sizeof(T) == 2) {
    LC_CHECK_CUDA(launch_kernel(qk_norm_rope_bf16_kernel, static_cast<unsigned>(ceil_div_ll(total / 2, 16)), 256, 0, s, 
        reinterpret_cast<bf16*>(qkv), ld, B, S, heads, eps, segs[0], s1, nseg));
    LC_LAUNCH_CHECK();
    return 0;
  }
  LC_CHECK_CUDA(launch_kernel(qk_norm_rope_kernel<T>, static_cast<unsigned>(ceil_div_ll(total, 8)), 256, 0, s, qkv, ld, B, S, heads, eps, segs[0],
                                                                                     s1, nseg));
  LC_LAUNCH_CHECK();
  return 0;
}

// [tokens, 128] fp32 cos / sin (each frequency repeated twice, interleaved: embeddings.py:315-327) -> [tokens, 64]
// half2 (cos, sin) per rotation pair for the fused qkv epilogue
__global__ void pack_rope_pairs_kernel(const float* __restrict__ cs, const float* __restrict__ sn, uint32_t* __restrict__ out,
                                       int n) {
  pdl_grid_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 64) return;
  const __half2 h = __floats2half2_rn(cs[2 * i], sn[2 * i]);
  out[i] = *reinterpret_cast<const uint32_t*>(&h);
}
int pack_rope_pairs(const float* cos, const float* sin, uint32_t* out, int n_tokens, cudaStream_t s) {
  LC_CHECK_CUDA(launch_kernel(pack_rope_pairs_kernel, ceil_div(n_tokens * 64, 256), 256, 0, s, cos, sin, out, n_tokens));
  LC_LAUNCH_CHECK();
  return 0;
}

template <typename T>
int patchify(const float* x, T* out, int B, int C, int THW, int Kp, cudaStream_t s) {
  dim3 grid(ceil_div(THW, 32), ceil_div(Kp, 32), B);
  ProfScope ps(PROF_MISC, 0.0, static_cast<double>(B) * THW * (C * 4.0 + Kp * sizeof(T)), s);
  LC_CHECK_CUDA(launch_kernel(patchify_kernel<T>, grid, 256, 0, s, x, out, C, THW, Kp));
  LC_LAUNCH_CHECK();
  return 0;
}

template <typename T>
int timestep_embed(const float* t, int n_t, int B, T* out, cudaStream_t s) {
  ProfScope ps(PROF_MISC, 0.0, static_cast<double>(B) * 256 * sizeof(T), s);
  LC_CHECK_CUDA(launch_kernel(timestep_embed_kernel<T>, ceil_div(B * 128, 128), 128, 0, s, t, n_t, B, out));
  LC_LAUNCH_CHECK();
  return 0;
}

template <typename T>
int token_mean(const float* x, int B, int N, int d, T* out, cudaStream_t s) {
  dim3 grid(ceil_div(d, 128), B);
  ProfScope ps(PROF_MISC, 0.0, static_cast<double>(B) * N * d * 4.0, s);
  LC_CHECK_CUDA(launch_kernel(token_mean_kernel<T>, grid, 256, 0, s, x, N, d, out));
  LC_LAUNCH_CHECK();
  return 0;
}

template <typename T>
int gated_add(float* h, const T* a, const float* gate, long long gate_stride, int M, int d, int rows_per_sample,
              cudaStream_t s) {
  const long long n4 = static_cast<long long>(M) * d / 4;
  ProfScope ps(PROF_MISC, 0.0, static_cast<double>(M) * d * (8.0 + sizeof(T)), s);
  LC_CHECK_CUDA(launch_kernel(gated_add_kernel<T>, static_cast<unsigned>(ceil_div_ll(n4, 256)), 256, 0, s, h, a, gate, gate_stride, n4, d,
                                                                                rows_per_sample));
  LC_LAUNCH_CHECK();
  return 0;
}

template <typename T>
int temb_combine(const float* a, const float* b, const float* sc, const float* sh, long long sc_stride, int B, int d,
                 float* out_f32, T* out_silu, cudaStream_t s) {
  ProfScope ps(PROF_MISC, 0.0, static_cast<double>(B) * d * 16.0, s);
  LC_CHECK_CUDA(launch_kernel(temb_combine_kernel<T>, ceil_div(B * d, 256), 256, 0, s, a, b, sc, sh, sc_stride, B, d, out_f32, out_silu));
  LC_LAUNCH_CHECK();
  return 0;
}

template <typename T>
int cast_rows(const float* x, T* out, long long n, cudaStream_t s) {
  ProfScope ps(PROF_MISC, 0.0, static_cast<double>(n) * (4.0 + sizeof(T)), s);
  LC_CHECK_CUDA(launch_kernel(cast_kernel<T>, static_cast<unsigned>(ceil_div_ll(n, 256)), 256, 0, s, x, out, n));
  LC_LAUNCH_CHECK();
  return 0;
}

int sched_dpmpp2m_step(const float* f, float* x, float* x0_prev, float* x_in_next, long long n, SchedCoef c,
                       cudaStream_t s) {
  LC_REQUIRE(n % 4 == 0, "scheduler: element count must be a multiple of 4");
  const long long n4 = n / 4;
  // algorithmic bytes: read F, x (+ previous x0 on 2M steps); write x0, x' (+ next x_in)
  ProfScope ps(PROF_SCHED, 0.0, static_cast<double>(n) * 4.0 * (4 + (c.a_d != 0.f ? 1 : 0) + (x_in_next != nullptr ? 1 : 0)), s);
  LC_CHECK_CUDA(launch_kernel(dpmpp2m_kernel, static_cast<unsigned>(ceil_div_ll(n4, 256)), 256, 0, s, f, x, x0_prev, x_in_next, n4, c));
  LC_LAUNCH_CHECK();
  return 0;
}

int sched_heun_step(const float* f, double* x, double* x_hat, double* d_cur, float* x_in_next, long long n, int phase,
                    double t_cur, double t_next, double c_skip, double c_out, double c_in_next, cudaStream_t s) {
  // algorithmic bytes: predictor reads F(4) x(8), writes x_hat d_cur x (24) + x_in(4); corrector reads F x x_hat d_cur
  // (28), writes x (8) + x_in (4)
  ProfScope ps(PROF_SCHED, 0.0, static_cast<double>(n) * (36.0 + (x_in_next != nullptr ? 4.0 : 0.0)), s);
  LC_CHECK_CUDA(launch_kernel(heun_kernel, static_cast<unsigned>(ceil_div_ll(n, 256)), 256, 0, s, f, x, x_hat, d_cur, x_in_next, n, phase, t_cur,
                                                                       t_next, c_skip, c_out, c_in_next));
  LC_LAUNCH_CHECK();
  return 0;
}

int sched_scale_input(const float* x, float* x_in, long long n, float c_in, cudaStream_t s) {
  LC_REQUIRE(n % 4 == 0, "scheduler: element count must be a multiple of 4");
  ProfScope ps(PROF_SCHED, 0.0, static_cast<double>(n) * 8.0, s);
  LC_CHECK_CUDA(launch_kernel(scale_kernel, static_cast<unsigned>(ceil_div_ll(n / 4, 256)), 256, 0, s, x, x_in, n / 4, c_in));
  LC_LAUNCH_CHECK();
  return 0;
}

int sched_heun_init(const float* noise, double* x, float* x_in, long long n, double t0, double c_in, cudaStream_t s) {
  ProfScope ps(PROF_SCHED, 0.0, static_cast<double>(n) * 16.0, s);
  LC_CHECK_CUDA(launch_kernel(heun_init_kernel, static_cast<unsigned>(ceil_div_ll(n, 256)), 256, 0, s, noise, x, x_in, n, t0, c_in));
  LC_LAUNCH_CHECK();
  return 0;
}

int sched_heun_churn(double* x, const double* noise, float* x_in, long long n, double k, double c_in, cudaStream_t s) {
  ProfScope ps(PROF_SCHED, 0.0, static_cast<double>(n) * 28.0, s);
  LC_CHECK_CUDA(launch_kernel(heun_churn_kernel, static_cast<unsigned>(ceil_div_ll(n, 256)), 256, 0, s, x, noise, x_in, n, k, c_in));
  LC_LAUNCH_CHECK();
  return 0;
}

int latent_feedback(const float* samples, float* known, float* phys, const float* mean, const float* stdv, float target,
                    int B, int C, int T, int t_in, int hw, cudaStream_t s) {
  LC_REQUIRE(hw % 2 == 0, "latent_feedback: h*w must be even");
  LC_REQUIRE(t_in >= 1 && t_in <= T, "latent_feedback: need 1 <= T_in <= T_out");
  LC_REQUIRE(phys == nullptr || (mean != nullptr && stdv != nullptr), "latent_feedback: mean/std required for the de-normalised output");
  const long long n2 = static_cast<long long>(B) * C * T * (hw / 2);
  ProfScope ps(PROF_SCHED, 0.0, static_cast<double>(n2) * 8.0 * (1.0 + (phys != nullptr ? 1.0 : 0.0) + static_cast<double>(t_in) / T), s);
  LC_CHECK_CUDA(launch_kernel(latent_feedback_kernel, static_cast<unsigned>(ceil_div_ll(n2, 256)), 256, 0, s, samples, known, phys, mean, stdv, target, C,
                                                                                   T, t_in, hw / 2, n2));
  LC_LAUNCH_CHECK();
  return 0;
}

#define LC_INST(T)                                                                                                    \
  template int layernorm_modulate<T>(const float*, T*, int, int, float, int, const float*, const float*, long long,   \
                                     const float*, const float*, cudaStream_t, int, int);                             \
  template int qk_norm_rope<T>(T*, long long, int, int, int, int, float, const RopeSeg*, int, cudaStream_t);          \
  template int patchify<T>(const float*, T*, int, int, int, int, cudaStream_t);                                       \
  template int timestep_embed<T>(const float*, int, int, T*, cudaStream_t);                                           \
  template int token_mean<T>(const float*, int, int, int, T*, cudaStream_t);                                          \
  template int gated_add<T>(float*, const T*, const float*, long long, int, int, int, cudaStream_t);                  \
  template int temb_combine<T>(const float*, const float*, const float*, const float*, long long, int, int, float*,   \
                               T*, cudaStream_t);                                                                     \
  template int cast_rows<T>(const float*, T*, long long, cudaStream_t);
LC_INST(float)
LC_INST(bf16)
#undef LC_INST

}  // namespace lc
