// Launch functions of the DC-AE decoder's non-GEMM kernels (NHWC internal layout).
#pragma once
#include "common.cuh"

namespace lc {
// Channel-major f32 source of the conv_in stage.  Frame f of a launch is global frame fg = frame0 + f = b * t_take + t
// and its channel-c plane starts at z + b * stride_b + c * stride_c + t * stride_t: a 4-D [n, C, H, W] batch
// (t_take = 1) or the first t_take frames of a 5-D [B, C, T, H, W] latent tensor read in place (decode_latent_ens,
// pipelines/utils.py:52-80, without the permute copy).  Optional per-channel de-normalisation
// (x / target) * scale[c] + shift[c] (inverse_normalize_transform_3D, dataloader/utils.py:233-240).
struct PlaneSrc {
  const float* z = nullptr;
  int frame0 = 0, t_take = 1;
  long long stride_b = 0, stride_c = 0, stride_t = 0;
  const float* scale = nullptr;
  const float* shift = nullptr;
  float target = 1.f;
  __host__ __device__ __forceinline__ long long plane(int f, int c) const {
    const int fg = frame0 + f;
    const int b = fg / t_take, t = fg - b * t_take;
    return b * stride_b + c * stride_c + t * stride_t;
  }
};
inline PlaneSrc plane_src_4d(const float* z, int C, int HW) {
  PlaneSrc p;
  p.z = z; p.stride_b = static_cast<long long>(C) * HW; p.stride_c = HW;
  return p;
}
template <typename T>
int pad_from_nchw(const PlaneSrc& z, T* out, int n, int C, int H, int W, int Cp, cudaStream_t s);
template <typename T>
int pad_from_nhwc(const float* x, T* out, int n, int C, int H, int W, int Cp, cudaStream_t s);
template <typename T>
int halo_fill(T* buf, int n, int H, int W, int Cp, cudaStream_t s);
template <typename T>
int dwconv3_glu(const T* in, const float* w, const float* bias, T* out, int n, int H, int W, int C, cudaStream_t s);
// T = activation storage type (bf16 production / float validation): q|k|v and the multiscale branch are stored as T
// (the reference's own bf16 autocast stores them as bf16 too; the attention core accumulates in fp32, DCAE.py:158-175)
template <typename T>
int multiscale_fused(const T* in, const float* w5, const float* wg, T* out, int n, int H, int W, int C, cudaStream_t s);
template <typename T>
int linear_attention(const T* qkv, const T* ms, T* out, int n, int HW, int heads, float eps, cudaStream_t s);
// y: [P, ldy] rows of type TY (a GEMM / conv output stored as T, or the fp32 residual stream)
template <typename TY, typename T>
int rmsnorm_rows(const TY* y, int ldy, const float* w, const float* b, float eps, float* resid, float* out_f32, T* out_t,
                 long long P, int C, int relu, cudaStream_t s, int pH = 0, int pW = 0, int pCp = 0);
template <typename TC, typename T>
int pixel_shuffle_shortcut(const TC* conv, const float* xin, float* out, T* out_t, int n, int H, int W, int Cin,
                           int Cout, cudaStream_t s, int pCp = 0);
// encoder: DCDownBlock2d tail (H x W = fine resolution) and the Encoder.forward output shortcut (+ normalisation)
template <typename T>
int pixel_unshuffle_shortcut(const float* conv, const float* xin, float* out, T* out_t, int n, int H, int W, int Cin,
                             int Cout, cudaStream_t s, int pCp = 0);
int enc_out_shortcut(float* out, const float* x, int n, int HW, int C, int L, const float* mean, const float* stdv,
                     float target, cudaStream_t s);
template <typename T>
int in_shortcut(float* x, T* x_t, const PlaneSrc& z, int n, int HW, int C, int Cz, cudaStream_t s);

// fp32 SIMT implicit 3x3 sphere conv on a padded NHWC f32 buffer (validation mode)
int conv3x3_f32(const float* xpad, int n_frames, int H, int W, int Cp, const float* wmat, int C_out,
                const EpiParams& epi, cudaStream_t stream);
}  // namespace lc
