// FP32 SIMT flash attention (validation mode): non-causal softmax(QK^T / sqrt(128)) V over a joint sequence.
// Reference: F.scaled_dot_product_attention call at LaDCast_3D_model.py:199-201 (no mask, default scale).
#include "kernels.h"

namespace lc {
namespace {

constexpr int HD = 128;
constexpr int BQ = 16;   // queries per CTA (4 per warp)
constexpr int BKV = 32;  // keys per tile (one per lane)
constexpr int KPAD = HD + 4;

__global__ void __launch_bounds__(128) attention_f32_kernel(const float* __restrict__ qkv, int S, int heads,
                                                            float* __restrict__ out_p, int Np,
                                                            float* __restrict__ out_c) {
  __shared__ __align__(16) float Qs[BQ][HD];
  __shared__ __align__(16) float Ks[BKV][KPAD];
  __shared__ __align__(16) float Vs[BKV][HD];
  __shared__ float Ps[4][4][BKV];
  const int d = heads * HD;
  const long long ld = 3LL * d;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * BQ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* base = qkv + static_cast<long long>(b) * S * ld;
  const float scale = 0.08838834764831845f;  // 1/sqrt(128)

  for (int i = threadIdx.x; i < BQ * (HD / 4); i += 128) {
    const int r = i / (HD / 4), c = (i % (HD / 4)) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + r < S) v = *reinterpret_cast<const float4*>(base + (q0 + r) * ld + h * HD + c);
    *reinterpret_cast<float4*>(&Qs[r][c]) = v;
  }
  float m[4], l[4];
  float4 o[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    m[q] = -INFINITY;
    l[q] = 0.f;
    o[q] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int k0 = 0; k0 < S; k0 += BKV) {
    __syncthreads();
    for (int i = threadIdx.x; i < BKV * (HD / 4); i += 128) {
      const int r = i / (HD / 4), c = (i % (HD / 4)) * 4;
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (k0 + r < S) {
        const float* rp = base + (k0 + r) * ld + h * HD + c;
        kv = *reinterpret_cast<const float4*>(rp + d);
        vv = *reinterpret_cast<const float4*>(rp + 2 * d);
      }
      *reinterpret_cast<float4*>(&Ks[r][c]) = kv;
      *reinterpret_cast<float4*>(&Vs[r][c]) = vv;
    }
    __syncthreads();
    float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
    for (int c = 0; c < HD; c += 4) {
      const float4 kk = *reinterpret_cast<const float4*>(&Ks[lane][c]);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 qq = *reinterpret_cast<const float4*>(&Qs[warp * 4 + q][c]);
        s[q] = fmaf(qq.x, kk.x, s[q]); s[q] = fmaf(qq.y, kk.y, s[q]);
        s[q] = fmaf(qq.z, kk.z, s[q]); s[q] = fmaf(qq.w, kk.w, s[q]);
      }
    }
    const bool valid = (k0 + lane) < S;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float sv = valid ? s[q] * scale : -INFINITY;
      float mx = sv;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      const float m_new = fmaxf(m[q], mx);
      const float p = valid ? expf(sv - m_new) : 0.f;
      float ps = p;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, off);
      const float alpha = (m[q] == -INFINITY) ? 0.f : expf(m[q] - m_new);
      l[q] = l[q] * alpha + ps;
      m[q] = m_new;
      o[q].x *= alpha; o[q].y *= alpha; o[q].z *= alpha; o[q].w *= alpha;
      Ps[warp][q][lane] = p;
    }
    __syncwarp();
#pragma unroll 4
    for (int j = 0; j < BKV; ++j) {
      const float4 vv = *reinterpret_cast<const float4*>(&Vs[j][lane * 4]);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float p = Ps[warp][q][j];
        o[q].x = fmaf(p, vv.x, o[q].x); o[q].y = fmaf(p, vv.y, o[q].y);
        o[q].z = fmaf(p, vv.z, o[q].z); o[q].w = fmaf(p, vv.w, o[q].w);
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int tok = q0 + warp * 4 + q;
    if (tok >= S) continue;
    const float inv = 1.0f / l[q];
    float* dst = (tok < Np) ? out_p + (static_cast<long long>(b) * Np + tok) * d
                            : out_c + (static_cast<long long>(b) * (S - Np) + (tok - Np)) * d;
    *reinterpret_cast<float4*>(dst + h * HD + lane * 4) =
        make_float4(o[q].x * inv, o[q].y * inv, o[q].z * inv, o[q].w * inv);
  }
}

}  // namespace

int attention_f32(const float* qkv, int B, int S, int heads, int head_dim, float* out_p, int Np, float* out_c,
                  cudaStream_t s) {
  LC_REQUIRE(head_dim == HD, "attention: head_dim must be 128");
  dim3 grid(ceil_div(S, BQ), heads, B);
  attention_f32_kernel<<<grid, 128, 0, s>>>(qkv, S, heads, out_p, Np, out_c);
  LC_LAUNCH_CHECK();
  return 0;
}

}  // namespace lc
