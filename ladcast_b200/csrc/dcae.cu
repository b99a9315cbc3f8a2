// DC-AE decoder handle: weight re-packing + Decoder.forward (reference models/DCAE.py:717-732 via
// AutoencoderDC.decode :1018-1056) as a fixed kernel sequence.  Internal layout is NHWC:
//   x    [n*H*W, C] f32  residual stream            padA/padB [n, H+2, W+2, Cp] T  sphere-padded conv inputs
//   y    [n*H*W, *] f32  raw conv / 1x1 outputs     xb [n*H*W, C] T                1x1-GEMM operand copy of x
// 3x3 sphere convolutions are implicit GEMMs on the tcgen05 kernel (gemm_tc.cu); 1x1 convolutions are plain
// GEMMs over pixels; depthwise/grouped/linear-attention/norm kernels are in dcae_kernels.cu.
#include <algorithm>
#include <cstdlib>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "../../include/ladcast_b200.h"
#include "dcae_kernels.h"
#include "gemm_tc.h"
#include "kernels.h"

namespace lc {
namespace {

struct Buf {
  void* p = nullptr;
  size_t bytes = 0;
  int alloc(size_t n) {
    if (n <= bytes) return 0;
    if (p) cudaFree(p);
    p = nullptr; bytes = 0;
    LC_CHECK_CUDA(cudaMalloc(&p, n));
    bytes = n;
    return 0;
  }
  void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
  template <typename U> U* as() const { return reinterpret_cast<U*>(p); }
};

struct St { float* p = nullptr; int64_t numel = 0; };

struct ConvW { void* w = nullptr; float* bias = nullptr; int cin = 0, cp = 0, cout = 0; };
struct MatW { void* w = nullptr; float* bias = nullptr; int out = 0, in = 0; };
struct EvitW {
  MatW qkv, to_out, inv, point;
  float *dw5 = nullptr, *g1 = nullptr, *no_w = nullptr, *no_b = nullptr, *dw3 = nullptr, *dw3_b = nullptr, *n_w = nullptr, *n_b = nullptr;
  int heads = 0, inner = 0;
};
struct ResW { ConvW c1, c2; float *n_w = nullptr, *n_b = nullptr; };
struct Block { int kind = 0; /*0 up, 1 res, 2 evit, 3 down*/ int cin = 0, cout = 0; ConvW up; ResW res; EvitW ev; };

inline int r64(int c) { return (c + 63) / 64 * 64; }
inline int r8(int c) { return (c + 7) / 8 * 8; }  // row pitch of T-typed GEMM outputs: 16-byte aligned bf16 rows

__global__ void pack_conv_kernel(const float* __restrict__ w, int cout, int cin, int cp, void* __restrict__ dst, int to_bf16) {
  // w [cout, cin, 3, 3] -> dst [cout, 9*cp], k = (ky*3+kx)*cp + c
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long K = 9LL * cp;
  if (i >= cout * K) return;
  const int o = static_cast<int>(i / K);
  const int k = static_cast<int>(i % K);
  const int tap = k / cp, c = k % cp;
  const float v = c < cin ? w[(static_cast<long long>(o) * cin + c) * 9 + tap] : 0.f;
  if (to_bf16) reinterpret_cast<bf16*>(dst)[i] = __float2bfloat16_rn(v);
  else reinterpret_cast<float*>(dst)[i] = v;
}

__global__ void transpose_taps_kernel(const float* __restrict__ w, int C, int taps, float* __restrict__ dst) {
  // depthwise weight [C, taps] -> tap-major [taps, C]
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C * taps) return;
  const int c = i / taps, t = i % taps;
  dst[static_cast<long long>(t) * C + c] = w[i];
}

__global__ void pack_mat_kernel(const float* __restrict__ w, long long n, void* __restrict__ dst, long long off, int to_bf16) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (to_bf16) reinterpret_cast<bf16*>(dst)[off + i] = __float2bfloat16_rn(w[i]);
  else reinterpret_cast<float*>(dst)[off + i] = w[i];
}

}  // namespace
}  // namespace lc

using namespace lc;

struct lc_dcae {
  lc_dcae_cfg cfg;
  bool f32 = false;
  size_t esz = 2;
  bool finalized = false;
  std::map<std::string, St> staged;
  std::vector<void*> owned;
  ConvW conv_in, conv_out;
  float *no_w = nullptr, *no_b = nullptr;
  std::vector<Block> blocks;
  bool has_decoder = false, has_encoder = false;
  bool fuse_norm = true;  // RMSNorm + residual in the conv / 1x1 epilogue where C <= 256 (LADCAST_B200_FUSE_NORM=0: off)
  ConvW enc_conv_in, enc_conv_out;
  std::vector<Block> enc_blocks;
  int max_frames = 0, h0 = 0, w0 = 0;
  Buf x, x2, y, padA, padB, xb, qkv, ms, att, hid, glu;
};

namespace lc {
namespace {

int find(lc_dcae* D, const std::string& k, const St** out) {
  auto it = D->staged.find(k);
  LC_REQUIRE(it != D->staged.end(), "missing checkpoint tensor '" + k + "'");
  *out = &it->second;
  return 0;
}
int fvec(lc_dcae* D, const std::string& k, int64_t n, float** out, cudaStream_t st, int64_t take = -1) {
  const St* s;
  LC_TRY(find(D, k, &s));
  LC_REQUIRE(s->numel == n, "unexpected size for '" + k + "'");
  const int64_t cnt = take > 0 ? take : n;
  LC_CHECK_CUDA(cudaMalloc(reinterpret_cast<void**>(out), static_cast<size_t>(cnt) * 4));
  D->owned.push_back(*out);
  LC_CHECK_CUDA(cudaMemcpyAsync(*out, s->p, static_cast<size_t>(cnt) * 4, cudaMemcpyDeviceToDevice, st));
  return 0;
}
int fvec_taps(lc_dcae* D, const std::string& k, int C, int taps, float** out, cudaStream_t st) {
  const St* s;
  LC_TRY(find(D, k, &s));
  LC_REQUIRE(s->numel == static_cast<int64_t>(C) * taps, "unexpected size for '" + k + "'");
  LC_CHECK_CUDA(cudaMalloc(reinterpret_cast<void**>(out), static_cast<size_t>(C) * taps * 4));
  D->owned.push_back(*out);
  transpose_taps_kernel<<<(C * taps + 255) / 256, 256, 0, st>>>(s->p, C, taps, *out);
  LC_LAUNCH_CHECK();
  return 0;
}
int make_conv(lc_dcae* D, const std::string& name, int cout_full, int cin, bool bias, ConvW* W, cudaStream_t st, int cout_keep = -1) {
  const St* w;
  LC_TRY(find(D, name + ".weight", &w));
  LC_REQUIRE(w->numel == static_cast<int64_t>(cout_full) * cin * 9, "unexpected conv weight shape for '" + name + "'");
  const int cout = cout_keep > 0 ? cout_keep : cout_full;
  W->cin = cin; W->cp = r64(cin); W->cout = cout;
  const long long n = static_cast<long long>(cout) * 9 * W->cp;
  LC_CHECK_CUDA(cudaMalloc(&W->w, static_cast<size_t>(n) * D->esz));
  D->owned.push_back(W->w);
  pack_conv_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(w->p, cout, cin, W->cp, W->w, D->f32 ? 0 : 1);
  LC_LAUNCH_CHECK();
  if (bias) LC_TRY(fvec(D, name + ".bias", cout_full, &W->bias, st, cout));
  return 0;
}
int make_mat(lc_dcae* D, const std::vector<std::string>& names, int in, bool bias, MatW* M, cudaStream_t st) {
  int total = 0;
  std::vector<const St*> ws;
  for (const auto& n : names) {
    const St* w;
    LC_TRY(find(D, n + ".weight", &w));
    LC_REQUIRE(w->numel % in == 0, "unexpected weight shape for '" + n + "'");
    ws.push_back(w);
    total += static_cast<int>(w->numel / in);
  }
  M->out = total; M->in = in;
  LC_CHECK_CUDA(cudaMalloc(&M->w, static_cast<size_t>(total) * in * D->esz));
  D->owned.push_back(M->w);
  long long off = 0;
  for (const St* w : ws) {
    pack_mat_kernel<<<static_cast<unsigned>((w->numel + 255) / 256), 256, 0, st>>>(w->p, w->numel, M->w, off, D->f32 ? 0 : 1);
    LC_LAUNCH_CHECK();
    off += w->numel;
  }
  if (bias) {
    LC_REQUIRE(names.size() == 1, "fused biased matrices unsupported");
    LC_TRY(fvec(D, names[0] + ".bias", total, &M->bias, st));
  }
  return 0;
}

template <typename T>
struct Run {
  lc_dcae* D;
  cudaStream_t st;
  int n;
  bool xb_valid = false;    // xb == T(x)?
  bool padA_valid = false;  // interior of padA == T(x) in the padded layout of the current (H, W, C)?

  int conv(const T* xpad, int H, int W, const ConvW& cw, const EpiParams& ep) const {
    if (sizeof(T) == 2) return conv3x3_bf16(xpad, n, H, W, cw.cp, cw.w, cw.cout, ep, st, cw.cin);
    return conv3x3_f32(reinterpret_cast<const float*>(xpad), n, H, W, cw.cp, reinterpret_cast<const float*>(cw.w), cw.cout, ep, st);
  }
  int gemm(const void* A, long long lda, long long M, const MatW& mw, void* out, bool out_f32, int act) const {
    GemmArgs g;
    g.A0 = A; g.lda0 = lda; g.K0 = mw.in; g.W = mw.w; g.ldw = mw.in; g.M = static_cast<int>(M); g.N = mw.out; g.K = mw.in;
    g.epi.mode = EPI_STORE; g.epi.act = act; g.epi.bias = mw.bias; g.epi.out = out; g.epi.ldo = mw.out;
    g.epi.out_f32 = (out_f32 || sizeof(T) == 4) ? 1 : 0;
    return sizeof(T) == 2 ? gemm_bf16(g, st) : gemm_f32(g, st);
  }
  EpiParams store_f32(float* out, int ld, const float* bias) const {
    EpiParams e;
    e.mode = EPI_STORE; e.out_f32 = 1; e.out = out; e.ldo = ld; e.bias = bias;
    return e;
  }
  // raw conv / 1x1 outputs that only feed a norm or a shuffle are stored as T (bf16 in production, like the
  // reference's own bf16 autocast), halving their write + read traffic
  EpiParams store_T(void* out, int ld, const float* bias) const {
    EpiParams e;
    e.mode = EPI_STORE; e.out_f32 = sizeof(T) == 4; e.out = out; e.ldo = ld; e.bias = bias;
    return e;
  }
  bool fused_norm_ok(int C) const { return sizeof(T) == 2 && D->fuse_norm && C <= 256 && C % 4 == 0; }
  EpiParams norm_epilogue(const float* nw, const float* nb, float eps, float* x, int C, void* out_t, int ldo) const {
    EpiParams e;
    e.mode = EPI_NORM_RESID; e.norm_w = nw; e.norm_b = nb; e.norm_eps = eps; e.xres = x; e.ldr = C; e.out = out_t; e.ldo = ldo;
    return e;
  }
  int gemm_epi(const void* A, long long lda, long long M, const MatW& mw, const EpiParams& e) const {
    GemmArgs g;
    g.A0 = A; g.lda0 = lda; g.K0 = mw.in; g.W = mw.w; g.ldw = mw.in; g.M = static_cast<int>(M); g.N = mw.out; g.K = mw.in;
    g.epi = e;
    return gemm_bf16(g, st);
  }
  int gemm_T(const void* A, long long lda, long long M, const MatW& mw, void* out, int ldo) const {
    GemmArgs g;
    g.A0 = A; g.lda0 = lda; g.K0 = mw.in; g.W = mw.w; g.ldw = mw.in; g.M = static_cast<int>(M); g.N = mw.out; g.K = mw.in;
    g.epi = store_T(out, ldo, mw.bias);
    return sizeof(T) == 2 ? gemm_bf16(g, st) : gemm_f32(g, st);
  }

  int res_block(const Block& b, int H, int W) {
    xb_valid = false;
    const ResW& w = b.res;
    const int C = b.cin;
    const long long P = static_cast<long long>(n) * H * W;
    T* padA = D->padA.as<T>();
    T* padB = D->padB.as<T>();
    float* x = D->x.as<float>();
    if (padA_valid) LC_TRY(halo_fill<T>(padA, n, H, W, w.c1.cp, st));  // the producer wrote the interior already
    else LC_TRY(pad_from_nhwc<T>(x, padA, n, C, H, W, w.c1.cp, st));
    padA_valid = false;
    // conv1 + bias + SiLU written straight into the interior of padB (conv2's padded input)
    EpiParams e;
    e.mode = EPI_STORE; e.act = ACT_SILU; e.bias = w.c1.bias; e.out = padB; e.ldo = w.c2.cp; e.out_f32 = sizeof(T) == 4;
    e.rows_per_sample = W; e.out_rows_per_sample = W + 2; e.out_row_offset = (W + 2) + 1;
    e.rows_per_group = H * W; e.group_extra_rows = 2 * (W + 2);
    LC_TRY(conv(padA, H, W, w.c1, e));
    LC_TRY(halo_fill<T>(padB, n, H, W, w.c2.cp, st));
    if (fused_norm_ok(C)) {
      // conv2 + RMSNorm + residual in ONE kernel: a 256-wide N tile holds the whole channel vector of a pixel, so the
      // epilogue normalises the accumulator rows in place, x += norm(y), and writes T(x) into padA's interior
      EpiParams f = norm_epilogue(w.n_w, w.n_b, 1e-5f, x, C, padA, w.c1.cp);
      f.rows_per_sample = W; f.out_rows_per_sample = W + 2; f.out_row_offset = (W + 2) + 1;
      f.rows_per_group = H * W; f.group_extra_rows = 2 * (W + 2);
      LC_TRY(conv(padB, H, W, w.c2, f));
    } else {
      LC_TRY(conv(padB, H, W, w.c2, store_T(D->y.p, r8(C), nullptr)));
      // x += norm(y); the new x is also written (as T) into padA's interior: the next 3x3 conv's padded input
      LC_TRY((rmsnorm_rows<T, T>(D->y.as<T>(), r8(C), w.n_w, w.n_b, 1e-5f, x, nullptr, padA, P, C, 0, st, H, W, w.c1.cp)));
    }
    padA_valid = true;
    return 0;
  }

  int evit_block(const Block& b, int H, int W) {
    padA_valid = false;
    const EvitW& w = b.ev;
    const int C = b.cin, HW = H * W;
    const long long P = static_cast<long long>(n) * HW;
    LC_REQUIRE(HW > D->cfg.head_dim, "quadratic-attention branch (H*W <= head_dim) is not implemented");
    float* x = D->x.as<float>();
    T* xb = D->xb.as<T>();
    T* y = D->y.as<T>();
    const int ldy = r8(C);
    if (!xb_valid) LC_TRY(cast_rows<T>(x, xb, P * C, st));
    xb_valid = true;
    // q|k|v and the multiscale branch are stored as T; the attention core accumulates in fp32 (DCAE.py:158-175)
    LC_TRY(gemm_T(xb, C, P, w.qkv, D->qkv.p, 3 * w.inner));
    LC_TRY(multiscale_fused<T>(D->qkv.as<T>(), w.dw5, w.g1, D->ms.as<T>(), n, H, W, 3 * w.inner, st));
    LC_TRY(linear_attention<T>(D->qkv.as<T>(), D->ms.as<T>(), D->att.as<T>(), n, HW, w.heads, 1e-15f, st));
    const bool fuse = fused_norm_ok(C);
    if (fuse) {
      LC_TRY(gemm_epi(D->att.p, 2 * w.inner, P, w.to_out, norm_epilogue(w.no_w, w.no_b, 1e-5f, x, C, xb, C)));
    } else {
      LC_TRY(gemm_T(D->att.p, 2 * w.inner, P, w.to_out, y, ldy));
      LC_TRY((rmsnorm_rows<T, T>(y, ldy, w.no_w, w.no_b, 1e-5f, x, nullptr, xb, P, C, 0, st)));
    }
    LC_TRY(gemm(xb, C, P, w.inv, D->hid.p, false, ACT_SILU));
    LC_TRY(dwconv3_glu<T>(D->hid.as<T>(), w.dw3, w.dw3_b, D->glu.as<T>(), n, H, W, 8 * C, st));
    if (fuse) return gemm_epi(D->glu.p, 4 * C, P, w.point, norm_epilogue(w.n_w, w.n_b, 1e-7f, x, C, xb, C));
    LC_TRY(gemm_T(D->glu.p, 4 * C, P, w.point, y, ldy));
    return rmsnorm_rows<T, T>(y, ldy, w.n_w, w.n_b, 1e-7f, x, nullptr, xb, P, C, 0, st);
  }

  int up_block(const Block& b, int& H, int& W, bool next_is_res) {
    const long long P = static_cast<long long>(n) * H * W;
    (void)P;
    if (padA_valid) LC_TRY(halo_fill<T>(D->padA.as<T>(), n, H, W, b.up.cp, st));
    else LC_TRY(pad_from_nhwc<T>(D->x.as<float>(), D->padA.as<T>(), n, b.cin, H, W, b.up.cp, st));
    padA_valid = false;
    LC_TRY(conv(D->padA.as<T>(), H, W, b.up, store_T(D->y.p, 4 * b.cout, b.up.bias)));
    // the shuffled result is also written as T: into padA's interior when a 3x3 conv follows, else as xb rows
    if (next_is_res) {
      LC_TRY((pixel_shuffle_shortcut<T, T>(D->y.as<T>(), D->x.as<float>(), D->x2.as<float>(), D->padA.as<T>(), n, H, W,
                                           b.cin, b.cout, st, r64(b.cout))));
      padA_valid = true;
      xb_valid = false;
    } else {
      LC_TRY((pixel_shuffle_shortcut<T, T>(D->y.as<T>(), D->x.as<float>(), D->x2.as<float>(), D->xb.as<T>(), n, H, W,
                                           b.cin, b.cout, st)));
      xb_valid = true;
    }
    std::swap(D->x, D->x2);
    H *= 2; W *= 2;
    return 0;
  }

  // DCDownBlock2d (DCAE.py:447-490): 3x3 conv C_in -> C_out/4 at the fine resolution, pixel_unshuffle(2), plus the
  // channel-averaged unshuffled input
  int down_block(const Block& b, int& H, int& W, bool next_is_res) {
    if (padA_valid) LC_TRY(halo_fill<T>(D->padA.as<T>(), n, H, W, b.up.cp, st));
    else LC_TRY(pad_from_nhwc<T>(D->x.as<float>(), D->padA.as<T>(), n, b.cin, H, W, b.up.cp, st));
    padA_valid = false;
    LC_TRY(conv(D->padA.as<T>(), H, W, b.up, store_f32(D->y.as<float>(), b.cout / 4, b.up.bias)));
    if (next_is_res) {
      LC_TRY(pixel_unshuffle_shortcut<T>(D->y.as<float>(), D->x.as<float>(), D->x2.as<float>(), D->padA.as<T>(), n, H, W,
                                         b.cin, b.cout, st, r64(b.cout)));
      padA_valid = true;
      xb_valid = false;
    } else {
      LC_TRY(pixel_unshuffle_shortcut<T>(D->y.as<float>(), D->x.as<float>(), D->x2.as<float>(), D->xb.as<T>(), n, H, W,
                                         b.cin, b.cout, st));
      xb_valid = true;
    }
    std::swap(D->x, D->x2);
    H /= 2; W /= 2;
    return 0;
  }

  // Encoder.forward (DCAE.py:617-631): x [n, in_channels, H, W] f32 NCHW -> out [n, latent, H/2^(ns-1), W/2^(ns-1)]
  int encode(const float* xin, int H, int W, float* out, const float* mean, const float* stdv, float target) {
    const lc_dcae_cfg& c = D->cfg;
    LC_TRY(pad_from_nchw<T>(plane_src_4d(xin, c.in_channels, H * W), D->padA.as<T>(), n, c.in_channels, H, W,
                            D->enc_conv_in.cp, st));
    LC_TRY(conv(D->padA.as<T>(), H, W, D->enc_conv_in,
                store_f32(D->x.as<float>(), D->enc_conv_in.cout, D->enc_conv_in.bias)));
    xb_valid = false;
    padA_valid = false;
    for (size_t bi = 0; bi < D->enc_blocks.size(); ++bi) {
      const Block& b = D->enc_blocks[bi];
      const bool next_is_res = bi + 1 < D->enc_blocks.size() && D->enc_blocks[bi + 1].kind == 1;
      if (b.kind == 3) LC_TRY(down_block(b, H, W, next_is_res));
      else if (b.kind == 1) LC_TRY(res_block(b, H, W));
      else LC_TRY(evit_block(b, H, W));
    }
    const int C = D->enc_conv_out.cin;
    if (padA_valid) LC_TRY(halo_fill<T>(D->padA.as<T>(), n, H, W, D->enc_conv_out.cp, st));
    else LC_TRY(pad_from_nhwc<T>(D->x.as<float>(), D->padA.as<T>(), n, C, H, W, D->enc_conv_out.cp, st));
    padA_valid = false;
    EpiParams e;
    e.mode = EPI_UNPATCHIFY; e.bias = D->enc_conv_out.bias; e.out = out; e.rows_per_sample = H * W;
    e.n_valid = c.latent_channels;
    LC_TRY(conv(D->padA.as<T>(), H, W, D->enc_conv_out, e));
    return enc_out_shortcut(out, D->x.as<float>(), n, H * W, C, c.latent_channels, mean, stdv, target, st);
  }

  // z: the n latent frames of this call (see PlaneSrc); out: [B, keep, out_T, 8h, 8w] with frame (z.frame0 + f) = b * out_T + t
  // written to plane (b, c, t) — out_T = 1 is the plain [n, keep, 8h, 8w] batch of AutoencoderDC.decode
  int decode(const PlaneSrc& z, int h, int w, float* out, int keep, const float* mean, const float* stdv, int out_T) {
    int H = h, W = w;
    const int C0 = D->conv_in.cout;
    LC_TRY(pad_from_nchw<T>(z, D->padA.as<T>(), n, D->cfg.latent_channels, H, W, D->conv_in.cp, st));
    LC_TRY(conv(D->padA.as<T>(), H, W, D->conv_in, store_f32(D->x.as<float>(), C0, D->conv_in.bias)));
    LC_TRY(in_shortcut<T>(D->x.as<float>(), D->xb.as<T>(), z, n, H * W, C0, D->cfg.latent_channels, st));
    xb_valid = true;
    for (size_t bi = 0; bi < D->blocks.size(); ++bi) {
      const Block& b = D->blocks[bi];
      const bool next_is_res = bi + 1 < D->blocks.size() && D->blocks[bi + 1].kind == 1;
      if (b.kind == 0) LC_TRY(up_block(b, H, W, next_is_res));
      else if (b.kind == 1) LC_TRY(res_block(b, H, W));
      else LC_TRY(evit_block(b, H, W));
    }
    const int C = D->conv_out.cin;
    const long long P = static_cast<long long>(n) * H * W;
    // norm_out + ReLU written straight into the padded input of conv_out
    LC_TRY((rmsnorm_rows<float, T>(D->x.as<float>(), C, D->no_w, D->no_b, 1e-7f, nullptr, nullptr, D->padA.as<T>(), P, C, 1,
                                   st, H, W, D->conv_out.cp)));
    LC_TRY(halo_fill<T>(D->padA.as<T>(), n, H, W, D->conv_out.cp, st));
    EpiParams e;
    e.mode = EPI_UNPATCHIFY; e.bias = D->conv_out.bias; e.out = out; e.rows_per_sample = H * W;
    e.n_valid = keep; e.ch_scale = stdv; e.ch_shift = mean;
    e.up_T = out_T; e.up_frame0 = out_T > 1 ? z.frame0 : 0;
    return conv(D->padA.as<T>(), H, W, D->conv_out, e);
  }
};

// ResBlock / EfficientViTBlock weights under key prefix p (same module classes in encoder and decoder, DCAE.py:417-444)
int build_block(lc_dcae* D, const std::string& p, bool evit, int C, Block* b, cudaStream_t st) {
  const lc_dcae_cfg& c = D->cfg;
  b->cin = b->cout = C;
  if (!evit) {
    b->kind = 1;
    LC_TRY(make_conv(D, p + ".conv1", C, C, true, &b->res.c1, st));
    LC_TRY(make_conv(D, p + ".conv2", C, C, false, &b->res.c2, st));
    LC_TRY(fvec(D, p + ".norm.weight", C, &b->res.n_w, st));
    LC_TRY(fvec(D, p + ".norm.bias", C, &b->res.n_b, st));
    return 0;
  }
  b->kind = 2;
  EvitW& e = b->ev;
  e.heads = C / c.head_dim;
  e.inner = e.heads * c.head_dim;
  LC_REQUIRE(c.head_dim == 32, "EfficientViT attention_head_dim must be 32");
  LC_TRY(make_mat(D, {p + ".attn.to_q", p + ".attn.to_k", p + ".attn.to_v"}, C, false, &e.qkv, st));
  LC_TRY(fvec_taps(D, p + ".attn.to_qkv_multiscale.0.proj_in.weight", 3 * e.inner, 25, &e.dw5, st));
  LC_TRY(fvec(D, p + ".attn.to_qkv_multiscale.0.proj_out.weight", 3LL * e.inner * 32, &e.g1, st));
  LC_TRY(make_mat(D, {p + ".attn.to_out"}, 2 * e.inner, false, &e.to_out, st));
  LC_TRY(fvec(D, p + ".attn.norm_out.weight", C, &e.no_w, st));
  LC_TRY(fvec(D, p + ".attn.norm_out.bias", C, &e.no_b, st));
  LC_TRY(make_mat(D, {p + ".conv_out.conv_inverted"}, C, true, &e.inv, st));
  LC_TRY(fvec_taps(D, p + ".conv_out.conv_depth.weight", 8 * C, 9, &e.dw3, st));
  LC_TRY(fvec(D, p + ".conv_out.conv_depth.bias", 8LL * C, &e.dw3_b, st));
  LC_TRY(make_mat(D, {p + ".conv_out.conv_point"}, 4 * C, false, &e.point, st));
  LC_TRY(fvec(D, p + ".conv_out.norm.weight", C, &e.n_w, st));
  LC_TRY(fvec(D, p + ".conv_out.norm.bias", C, &e.n_b, st));
  return 0;
}

int finalize_decoder(lc_dcae* D, cudaStream_t st) {
  const lc_dcae_cfg& c = D->cfg;
  const int ns = c.n_stages;
  const int Ctop = c.stage_channels[ns - 1];
  LC_TRY(make_conv(D, "decoder.conv_in", Ctop, c.latent_channels, true, &D->conv_in, st));
  int j = 0;
  for (int i = ns - 1; i >= 0; --i) {
    const int C = c.stage_channels[i];
    if (i < ns - 1 && c.stage_layers[i] > 0) {
      Block b;
      b.kind = 0; b.cin = c.stage_channels[i + 1]; b.cout = C;
      LC_REQUIRE((4 * b.cout) % b.cin == 0, "up-block shortcut needs 4*C_out divisible by C_in");
      LC_TRY(make_conv(D, "decoder.up_blocks." + std::to_string(j) + ".conv", 4 * C, b.cin, true, &b.up, st));
      D->blocks.push_back(b);
      ++j;
    }
    for (int l = 0; l < c.stage_layers[i]; ++l) {
      Block b;
      LC_TRY(build_block(D, "decoder.up_blocks." + std::to_string(j), c.stage_is_evit[i] != 0, C, &b, st));
      D->blocks.push_back(b);
      ++j;
    }
  }
  const int C0 = c.stage_channels[0];
  LC_TRY(fvec(D, "decoder.norm_out.weight", C0, &D->no_w, st));
  LC_TRY(fvec(D, "decoder.norm_out.bias", C0, &D->no_b, st));
  LC_TRY(make_conv(D, "decoder.conv_out", c.out_channels, C0, true, &D->conv_out, st));
  D->has_decoder = true;
  return 0;
}

// Encoder (DCAE.py:539-615), layers_per_block[0] > 0 variant: conv_in is a plain SphereConv2d
int finalize_encoder(lc_dcae* D, cudaStream_t st) {
  const lc_dcae_cfg& c = D->cfg;
  const int ns = c.n_stages;
  LC_REQUIRE(c.in_channels > 0, "encoder weights given but lc_dcae_cfg.in_channels is not set");
  LC_REQUIRE(c.enc_stage_layers[0] > 0, "encoder with layers_per_block[0] == 0 (down-sampling conv_in) is not implemented");
  LC_REQUIRE(c.enc_stage_channels[ns - 1] % c.latent_channels == 0, "encoder out shortcut needs C_top divisible by latent_channels");
  LC_TRY(make_conv(D, "encoder.conv_in", c.enc_stage_channels[0], c.in_channels, true, &D->enc_conv_in, st));
  int j = 0;
  for (int i = 0; i < ns; ++i) {
    const int C = c.enc_stage_channels[i];
    for (int l = 0; l < c.enc_stage_layers[i]; ++l) {
      Block b;
      LC_TRY(build_block(D, "encoder.down_blocks." + std::to_string(j), c.enc_stage_is_evit[i] != 0, C, &b, st));
      D->enc_blocks.push_back(b);
      ++j;
    }
    if (i < ns - 1 && c.enc_stage_layers[i] > 0) {
      Block b;
      b.kind = 3; b.cin = C; b.cout = c.enc_stage_channels[i + 1];
      LC_REQUIRE(b.cout % 4 == 0 && (4 * b.cin) % b.cout == 0, "down-block needs C_out % 4 == 0 and 4*C_in divisible by C_out");
      LC_TRY(make_conv(D, "encoder.down_blocks." + std::to_string(j) + ".conv", b.cout / 4, b.cin, true, &b.up, st));
      D->enc_blocks.push_back(b);
      ++j;
    }
  }
  LC_TRY(make_conv(D, "encoder.conv_out", c.latent_channels, c.enc_stage_channels[ns - 1], true, &D->enc_conv_out, st));
  D->has_encoder = true;
  return 0;
}

int finalize_impl(lc_dcae* D, cudaStream_t st) {
  const bool dec = D->staged.count("decoder.conv_in.weight") != 0, enc = D->staged.count("encoder.conv_in.weight") != 0;
  LC_REQUIRE(dec || enc, "no decoder.* or encoder.* tensors were loaded");
  if (dec) LC_TRY(finalize_decoder(D, st));
  if (enc) LC_TRY(finalize_encoder(D, st));
  LC_CHECK_CUDA(cudaStreamSynchronize(st));
  for (auto& kv : D->staged) cudaFree(kv.second.p);
  D->staged.clear();
  D->finalized = true;
  return 0;
}

}  // namespace
}  // namespace lc

extern "C" {

int lc_dcae_create(const lc_dcae_cfg* cfg, lc_dcae** out) {
  LC_REQUIRE(cfg && out, "null argument");
  LC_REQUIRE(cfg->n_stages >= 1 && cfg->n_stages <= 8, "n_stages out of range");
  LC_REQUIRE(cfg->out_channels > 0 && cfg->out_channels <= 128, "decoder out_channels must be <= 128");
  if (cfg->precision == LC_PRECISION_BF16)
    for (int i = 0; i < cfg->n_stages; ++i)
      LC_REQUIRE(!cfg->stage_is_evit[i] || cfg->stage_layers[i] == 0 || cfg->stage_channels[i] % 8 == 0,
                 "bf16 decoder needs EfficientViT stage channels divisible by 8 (TMA 16-byte row pitch)");
  for (int i = 0; i < cfg->n_stages; ++i) LC_REQUIRE(cfg->stage_channels[i] % 4 == 0, "stage channels must be multiples of 4");
  if (cfg->in_channels > 0)
    for (int i = 0; i < cfg->n_stages; ++i) {
      LC_REQUIRE(cfg->enc_stage_channels[i] > 0 && cfg->enc_stage_channels[i] % 4 == 0, "encoder stage channels must be multiples of 4");
      LC_REQUIRE(cfg->precision != LC_PRECISION_BF16 || !cfg->enc_stage_is_evit[i] || cfg->enc_stage_layers[i] == 0 ||
                     cfg->enc_stage_channels[i] % 8 == 0,
                 "bf16 encoder needs EfficientViT stage channels divisible by 8 (TMA 16-byte row pitch)");
    }
  LC_REQUIRE(cfg->stage_channels[cfg->n_stages - 1] % cfg->latent_channels == 0, "in_shortcut needs C_top divisible by latent_channels");
  lc_dcae* D = new lc_dcae();
  D->cfg = *cfg;
  D->f32 = cfg->precision == LC_PRECISION_F32;
  D->esz = D->f32 ? 4 : 2;
  const char* fn = getenv("LADCAST_B200_FUSE_NORM");
  D->fuse_norm = !(fn != nullptr && fn[0] == '0');
  *out = D;
  return 0;
}

void lc_dcae_destroy(lc_dcae* D) {
  if (!D) return;
  for (void* p : D->owned) cudaFree(p);
  for (auto& kv : D->staged) cudaFree(kv.second.p);
  Buf* bufs[] = {&D->x, &D->x2, &D->y, &D->padA, &D->padB, &D->xb, &D->qkv, &D->ms, &D->att, &D->hid, &D->glu};
  for (Buf* b : bufs) b->release();
  delete D;
}

int lc_dcae_load(lc_dcae* D, const char* key, const float* data, const int64_t* shape, int ndim, void* stream) {
  LC_REQUIRE(D && key && data && shape, "null argument");
  LC_REQUIRE(!D->finalized, "lc_dcae_load after finalize");
  St s;
  s.numel = 1;
  for (int i = 0; i < ndim; ++i) s.numel *= shape[i];
  LC_CHECK_CUDA(cudaMalloc(reinterpret_cast<void**>(&s.p), static_cast<size_t>(s.numel) * 4));
  LC_CHECK_CUDA(cudaMemcpyAsync(s.p, data, static_cast<size_t>(s.numel) * 4, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
  auto it = D->staged.find(key);
  if (it != D->staged.end()) cudaFree(it->second.p);
  D->staged[key] = s;
  return 0;
}

int lc_dcae_finalize(lc_dcae* D, void* stream) {
  LC_REQUIRE(D && !D->finalized, "finalize called twice or on null handle");
  return finalize_impl(D, static_cast<cudaStream_t>(stream));
}

int lc_dcae_reserve(lc_dcae* D, int max_frames, int h, int w, void* stream) {
  LC_REQUIRE(D && D->finalized, "reserve before finalize");
  LC_REQUIRE(max_frames > 0 && h > 0 && w > 0 && w % 2 == 0, "bad decoder geometry (width must be even)");
  const lc_dcae_cfg& c = D->cfg;
  const int ns = c.n_stages;
  size_t mx_x = 0, mx_y = 0, mx_pad = 0, mx_xb = 0, mx_qkv = 0, mx_att = 0, mx_hid = 0, mx_glu = 0;
  // stage i (0 = highest resolution) has spatial size (h, w) << (ns-1-i)
  for (int i = ns - 1; i >= 0; --i) {
    const size_t H = static_cast<size_t>(h) << (ns - 1 - i), W = static_cast<size_t>(w) << (ns - 1 - i);
    const size_t C = c.stage_channels[i], P = H * W;
    mx_x = std::max(mx_x, P * C);
    mx_y = std::max(mx_y, P * static_cast<size_t>(r8(static_cast<int>(C))));
    mx_pad = std::max(mx_pad, (H + 2) * (W + 2) * static_cast<size_t>(r64(static_cast<int>(C))));
    if (i < ns - 1) {  // up-block conv runs at the coarser resolution, producing 4*C channels
      const size_t Hc = H / 2, Wc = W / 2, Cc = c.stage_channels[i + 1];
      mx_y = std::max(mx_y, Hc * Wc * 4 * C);
      mx_pad = std::max(mx_pad, (Hc + 2) * (Wc + 2) * static_cast<size_t>(r64(static_cast<int>(Cc))));
    }
    if (c.stage_is_evit[i] && c.stage_layers[i] > 0) {
      const size_t inner = (C / c.head_dim) * c.head_dim;
      mx_qkv = std::max(mx_qkv, P * 3 * inner);
      mx_att = std::max(mx_att, P * 2 * inner);
      mx_hid = std::max(mx_hid, P * 8 * C);
      mx_glu = std::max(mx_glu, P * 4 * C);
    }
    mx_xb = std::max(mx_xb, P * C);
  }
  {
    const size_t H = h, W = w;
    mx_pad = std::max(mx_pad, (H + 2) * (W + 2) * static_cast<size_t>(r64(c.latent_channels)));
  }
  if (c.in_channels > 0) {  // encoder stages (same resolutions, possibly other widths) + its 89-channel input
    for (int i = ns - 1; i >= 0; --i) {
      const size_t H = static_cast<size_t>(h) << (ns - 1 - i), W = static_cast<size_t>(w) << (ns - 1 - i);
      const size_t C = c.enc_stage_channels[i], P = H * W;
      mx_x = std::max(mx_x, P * C);
      mx_y = std::max(mx_y, P * static_cast<size_t>(r8(static_cast<int>(C))));
      mx_xb = std::max(mx_xb, P * C);
      mx_pad = std::max(mx_pad, (H + 2) * (W + 2) * static_cast<size_t>(r64(static_cast<int>(C))));
      if (c.enc_stage_is_evit[i] && c.enc_stage_layers[i] > 0) {
        const size_t inner = (C / c.head_dim) * c.head_dim;
        mx_qkv = std::max(mx_qkv, P * 3 * inner);
        mx_att = std::max(mx_att, P * 2 * inner);
        mx_hid = std::max(mx_hid, P * 8 * C);
        mx_glu = std::max(mx_glu, P * 4 * C);
      }
    }
    const size_t H = static_cast<size_t>(h) << (ns - 1), W = static_cast<size_t>(w) << (ns - 1);
    mx_pad = std::max(mx_pad, (H + 2) * (W + 2) * static_cast<size_t>(r64(c.in_channels)));
  }
  const size_t n = max_frames, e = D->esz;
  LC_TRY(D->x.alloc(n * mx_x * 4)); LC_TRY(D->x2.alloc(n * mx_x * 4)); LC_TRY(D->y.alloc(n * mx_y * 4));
  LC_TRY(D->padA.alloc(n * mx_pad * e)); LC_TRY(D->padB.alloc(n * mx_pad * e));
  LC_TRY(D->xb.alloc(n * mx_xb * e));
  LC_TRY(D->qkv.alloc(n * std::max<size_t>(mx_qkv, 1) * 4));
  LC_TRY(D->ms.alloc(n * std::max<size_t>(mx_qkv, 1) * 4));
  LC_TRY(D->att.alloc(n * std::max<size_t>(mx_att, 1) * e)); LC_TRY(D->hid.alloc(n * std::max<size_t>(mx_hid, 1) * e));
  LC_TRY(D->glu.alloc(n * std::max<size_t>(mx_glu, 1) * e));
  // padB's channel padding is never written by a conv epilogue: keep it zero
  LC_CHECK_CUDA(cudaMemsetAsync(D->padB.p, 0, D->padB.bytes, static_cast<cudaStream_t>(stream)));
  LC_CHECK_CUDA(cudaMemsetAsync(D->padA.p, 0, D->padA.bytes, static_cast<cudaStream_t>(stream)));
  D->max_frames = max_frames; D->h0 = h; D->w0 = w;
  return 0;
}

int lc_dcae_decode(lc_dcae* D, const float* z, int n, int h, int w, float* out, int keep_channels, const float* mean,
                   const float* stdv, void* stream) {
  LC_REQUIRE(D && D->max_frames > 0, "decode before reserve");
  LC_REQUIRE(D->has_decoder, "no decoder.* weights were loaded into this handle");
  LC_REQUIRE(n > 0 && n <= D->max_frames && h == D->h0 && w == D->w0, "decode geometry differs from lc_dcae_reserve");
  LC_REQUIRE(z && out, "null argument");
  LC_REQUIRE(keep_channels > 0 && keep_channels <= D->cfg.out_channels, "keep_channels out of range");
  LC_REQUIRE((mean == nullptr) == (stdv == nullptr), "mean and std must be given together");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // the halo/channel padding of padB must be zero where conv epilogues do not write; re-zero is not needed between
  // calls because epilogues only ever write real channels and halo_fill rewrites the halo.
  const PlaneSrc src = plane_src_4d(z, D->cfg.latent_channels, h * w);
  if (D->f32) {
    Run<float> r;
    r.D = D; r.st = st; r.n = n;
    return r.decode(src, h, w, out, keep_channels, mean, stdv, 1);
  }
  Run<bf16> r;
  r.D = D; r.st = st; r.n = n;
  return r.decode(src, h, w, out, keep_channels, mean, stdv, 1);
}

int lc_dcae_decode_ens(lc_dcae* D, const float* latents, int batch, int t_total, int t_take, int frame0, int n, int h, int w,
                       float* out, int keep_channels, const float* mean, const float* stdv, const float* lat_mean,
                       const float* lat_std, float target_std, void* stream) {
  LC_REQUIRE(D && D->max_frames > 0, "decode before reserve");
  LC_REQUIRE(D->has_decoder, "no decoder.* weights were loaded into this handle");
  LC_REQUIRE(latents && out, "null argument");
  LC_REQUIRE(batch > 0 && t_total > 0 && t_take > 0 && t_take <= t_total, "bad (batch, T, extract_first)");
  LC_REQUIRE(frame0 >= 0 && n > 0 && frame0 + n <= batch * t_take, "frame range outside the [batch, extract_first] block");
  LC_REQUIRE(n <= D->max_frames && h == D->h0 && w == D->w0, "decode geometry differs from lc_dcae_reserve");
  LC_REQUIRE(keep_channels > 0 && keep_channels <= D->cfg.out_channels, "keep_channels out of range");
  LC_REQUIRE((mean == nullptr) == (stdv == nullptr), "mean and std must be given together");
  LC_REQUIRE((lat_mean == nullptr) == (lat_std == nullptr), "latent mean and std must be given together");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PlaneSrc src;
  src.z = latents; src.frame0 = frame0; src.t_take = t_take;
  src.stride_t = static_cast<long long>(h) * w;
  src.stride_c = src.stride_t * t_total;
  src.stride_b = src.stride_c * D->cfg.latent_channels;
  src.scale = lat_std; src.shift = lat_mean; src.target = target_std;
  if (D->f32) {
    Run<float> r;
    r.D = D; r.st = st; r.n = n;
    return r.decode(src, h, w, out, keep_channels, mean, stdv, t_take);
  }
  Run<bf16> r;
  r.D = D; r.st = st; r.n = n;
  return r.decode(src, h, w, out, keep_channels, mean, stdv, t_take);
}

int lc_dcae_encode(lc_dcae* D, const float* x, int n, int height, int width, float* out, const float* mean,
                   const float* stdv, float target_std, void* stream) {
  LC_REQUIRE(D && D->max_frames > 0, "encode before reserve");
  LC_REQUIRE(D->has_encoder, "no encoder.* weights were loaded into this handle");
  const int r = 1 << (D->cfg.n_stages - 1);
  LC_REQUIRE(n > 0 && n <= D->max_frames && height == D->h0 * r && width == D->w0 * r,
             "encode geometry differs from lc_dcae_reserve (fields must be 2^(n_stages-1) x the reserved latent size)");
  LC_REQUIRE(x && out, "null argument");
  LC_REQUIRE((mean == nullptr) == (stdv == nullptr), "mean and std must be given together");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (D->f32) {
    Run<float> rr;
    rr.D = D; rr.st = st; rr.n = n;
    return rr.encode(x, height, width, out, mean, stdv, target_std);
  }
  Run<bf16> rr;
  rr.D = D; rr.st = st; rr.n = n;
  return rr.encode(x, height, width, out, mean, stdv, target_std);
}

// Test exports of the decoder / encoder index permutations (NHWC f32 in and out, fp32 arithmetic: one add per element,
// so the result is bit-identical to torch's pixel_shuffle / pixel_unshuffle formulation, DCAE.py:476-490, 519-536).
int lc_pixel_shuffle_shortcut(const float* conv, const float* xin, float* out, int n, int H, int W, int cin, int cout,
                              void* stream) {
  LC_REQUIRE(conv && xin && out, "null argument");
  return pixel_shuffle_shortcut<float, float>(conv, xin, out, nullptr, n, H, W, cin, cout, static_cast<cudaStream_t>(stream));
}
int lc_pixel_unshuffle_shortcut(const float* conv, const float* xin, float* out, int n, int H, int W, int cin, int cout,
                                void* stream) {
  LC_REQUIRE(conv && xin && out, "null argument");
  return pixel_unshuffle_shortcut<float>(conv, xin, out, nullptr, n, H, W, cin, cout, static_cast<cudaStream_t>(stream));
}

// Test export: one 3x3 sphere convolution (+bias, act) NCHW f32 -> NCHW f32 through the implicit-GEMM path.
int lc_sphere_conv3x3(int precision, const float* x, const float* w, const float* bias, float* out, int n, int cin, int H,
                      int W, int cout, int act, void* stream) {
  LC_REQUIRE(x && w && out, "null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool f32 = precision == LC_PRECISION_F32;
  const size_t e = f32 ? 4 : 2;
  const int cp = r64(cin);
  void *xpad = nullptr, *wm = nullptr;
  LC_CHECK_CUDA(cudaMalloc(&xpad, static_cast<size_t>(n) * (H + 2) * (W + 2) * cp * e));
  LC_CHECK_CUDA(cudaMalloc(&wm, static_cast<size_t>(cout) * 9 * cp * e));
  const long long nw = static_cast<long long>(cout) * 9 * cp;
  pack_conv_kernel<<<static_cast<unsigned>((nw + 255) / 256), 256, 0, st>>>(w, cout, cin, cp, wm, f32 ? 0 : 1);
  EpiParams ep;
  ep.mode = EPI_UNPATCHIFY; ep.act = act; ep.bias = bias; ep.out = out; ep.rows_per_sample = H * W; ep.n_valid = cout;
  int rc;
  if (f32) {
    rc = pad_from_nchw<float>(plane_src_4d(x, cin, H * W), reinterpret_cast<float*>(xpad), n, cin, H, W, cp, st);
    if (rc == 0) rc = conv3x3_f32(reinterpret_cast<float*>(xpad), n, H, W, cp, reinterpret_cast<float*>(wm), cout, ep, st);
  } else {
    rc = pad_from_nchw<bf16>(plane_src_4d(x, cin, H * W), reinterpret_cast<bf16*>(xpad), n, cin, H, W, cp, st);
    if (rc == 0) rc = conv3x3_bf16(xpad, n, H, W, cp, wm, cout, ep, st);
  }
  cudaStreamSynchronize(st);
  cudaFree(xpad);
  cudaFree(wm);
  return rc;
}

}  // extern "C"
