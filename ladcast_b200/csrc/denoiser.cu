// Denoiser handle: weight re-packing + the forward pass of LaDCastTransformer3DModel
// (reference: models/LaDCast_3D_model.py:833-1071) as a fixed sequence of library kernels on one stream.
//
// Data layout in HBM (B = members, Np = T_out*H*W pred tokens, Nc = T_in*H*W cond tokens, S = Np+Nc, d = hidden):
//   h   [B*Np, d] f32   pred residual stream            e   [B*Nc, d] f32   cond residual stream
//   n_* [rows, d]  T    LayerNorm+modulate outputs      qkv [B, S, 3d] T    joint q|k|v (pred tokens first)
//   att_* [rows, d] T   attention output, split         mlp_* [rows, 4d] T  MLP hidden
//   mod [B, 12d*n_dual + 3d*n_single + 2d] f32          all AdaLN modulation vectors from ONE GEMM per call
// T = bf16 (tensor-core mode) or float (FP32 validation mode).  The residual streams stay fp32 in both modes.
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

#include "../../include/ladcast_b200.h"
#include "kernels.h"

namespace lc {

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int alloc(size_t n) {
    if (n <= bytes) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    LC_CHECK_CUDA(cudaMalloc(&p, n));
    bytes = n;
    return 0;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  template <typename U>
  U* as() const { return reinterpret_cast<U*>(p); }
};

struct Lin {
  void* w = nullptr;      // [out, ldw] storage type T
  float* bias = nullptr;  // [out] or null
  int out = 0, in = 0, ldw = 0;
};

struct Staged {
  float* p = nullptr;
  std::vector<int64_t> shape;
  int64_t numel = 0;
};

struct RefinerW {
  float *n1w, *n1b, *n2w, *n2b, *nq, *nk;
  Lin qkv, ff0, ff2;
};
struct DualW {
  Lin qkv, add_qkv, to_out, to_add_out, ff0, ff2, ffc0, ffc2;
  float *nq, *nk, *naq, *nak;
};
struct SingleW {
  Lin qkv, mlp, proj_out;
  float *nq, *nk;
};

}  // namespace lc

using namespace lc;

struct lc_denoiser {
  lc_denoiser_cfg cfg;
  bool f32 = false;
  int d = 0;
  size_t esz = 2;  // sizeof(T)
  bool finalized = false;
  std::map<std::string, Staged> staged;
  std::vector<void*> owned;  // weight allocations

  Lin x_emb, c_emb, r_t1, r_t2, r_p1, r_p2, r_proj_in, r_gates, t1, t2, p1, p2, te1, te2, mod, proj_out;
  std::vector<RefinerW> refiner;
  std::vector<DualW> dual;
  std::vector<SingleW> single;
  int mod_dim = 0, kp_in = 96;
  // RMSNorm(q,k)+RoPE fused into the qkv GEMM epilogue (gemm_tc.cu K_QKV: two passes over tensor memory, staged
  // coalesced stores; a one-pass packed-bf16 form measured the same).  Measured in-step A/B on B200 (375M, B=20), three epilogue forms over
  // two rounds: always +0.6 % (477.5 vs 480.5 ms per AR step): the norm / rotation arithmetic of a 128x256 tile on 8
  // epilogue warps (~2x the plain epilogue) exceeds the K=1536 main loop it has to hide under, and the qkv GEMMs give
  // back the 18 ms of the removed HBM-bound kernel (GEMM class 256 -> 275 ms).  Off by default; LADCAST_B200_FUSE_QK=1.
  bool fuse_qk = false;
  // single-stream blocks: pred and cond rows in one GEMM launch per projection (LADCAST_B200_MERGE_STREAMS=0: off)
  bool merge_streams = true;

  // geometry
  int maxB = 0, T_in = 0, T_out = 0, H = 0, W = 0, Np = 0, Nc = 0, S = 0;
  int curB = 0, n_ts = 0;
  DevBuf cos_p, sin_p, cos_c, sin_c, cs_p, cs_c;  // cs_*: [tokens, 64] half2 (cos, sin) for the fused qkv epilogue
  // workspace
  // h / n_p / att_p / mlp_p hold the pred-token rows followed by the cond-token rows of the current batch
  DevBuf tok_x, tok_c, h, e0, e0T, e_proj, n_p, qkv, att_p, mlp_p;
  DevBuf sincos, tmpA, tmpB, r_te, r_pe, r_tembS, gates, t_te, pooled, pe, temb, tembS, modv, te_out, yearT;
  float* e_ptr() const { return h.as<float>() + static_cast<size_t>(curB) * Np * d; }  // cond-token residual stream
};

namespace lc {
namespace {

__global__ void pack_rows_kernel(const float* __restrict__ src, int rows, int in, void* __restrict__ dst, int ldw,
                                 int row_off, int to_bf16) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(rows) * ldw) return;
  const int r = static_cast<int>(i / ldw), c = static_cast<int>(i % ldw);
  const float v = c < in ? src[static_cast<long long>(r) * in + c] : 0.f;
  const long long o = static_cast<long long>(row_off + r) * ldw + c;
  if (to_bf16) reinterpret_cast<bf16*>(dst)[o] = __float2bfloat16_rn(v);
  else reinterpret_cast<float*>(dst)[o] = v;
}

int find(lc_denoiser* D, const std::string& key, const Staged** out) {
  auto it = D->staged.find(key);
  LC_REQUIRE(it != D->staged.end(), "missing checkpoint tensor '" + key + "'");
  *out = &it->second;
  return 0;
}

// Fuse the listed Linear layers (rows concatenated) into one [sum(out), ldw] matrix of the storage type.
int make_lin(lc_denoiser* D, const std::vector<std::string>& names, int in, int ldw, bool has_bias, Lin* L,
             cudaStream_t st) {
  int total = 0;
  std::vector<const Staged*> ws, bs;
  for (const auto& n : names) {
    const Staged* w;
    LC_TRY(find(D, n + ".weight", &w));
    LC_REQUIRE(w->numel % in == 0, "unexpected weight shape for '" + n + "'");
    ws.push_back(w);
    total += static_cast<int>(w->numel / in);
    if (has_bias) {
      const Staged* b;
      LC_TRY(find(D, n + ".bias", &b));
      bs.push_back(b);
    }
  }
  L->out = total;
  L->in = in;
  L->ldw = ldw;
  void* wbuf = nullptr;
  LC_CHECK_CUDA(cudaMalloc(&wbuf, static_cast<size_t>(total) * ldw * D->esz));
  D->owned.push_back(wbuf);
  L->w = wbuf;
  if (has_bias) {
    LC_CHECK_CUDA(cudaMalloc(reinterpret_cast<void**>(&L->bias), static_cast<size_t>(total) * 4));
    D->owned.push_back(L->bias);
  }
  int off = 0;
  for (size_t i = 0; i < ws.size(); ++i) {
    const int rows = static_cast<int>(ws[i]->numel / in);
    const long long n = static_cast<long long>(rows) * ldw;
    pack_rows_kernel<<<static_cast<unsigned>(ceil_div_ll(n, 256)), 256, 0, st>>>(ws[i]->p, rows, in, wbuf, ldw, off,
                                                                               D->f32 ? 0 : 1);
    LC_LAUNCH_CHECK();
    if (has_bias) {
      LC_REQUIRE(bs[i]->numel == rows, "bias shape mismatch for '" + names[i] + "'");
      LC_CHECK_CUDA(cudaMemcpyAsync(L->bias + off, bs[i]->p, static_cast<size_t>(rows) * 4, cudaMemcpyDeviceToDevice, st));
    }
    off += rows;
  }
  return 0;
}

int vec(lc_denoiser* D, const std::string& key, int n, float** out, cudaStream_t st) {
  const Staged* s;
  LC_TRY(find(D, key, &s));
  LC_REQUIRE(s->numel == n, "unexpected vector length for '" + key + "'");
  LC_CHECK_CUDA(cudaMalloc(reinterpret_cast<void**>(out), static_cast<size_t>(n) * 4));
  D->owned.push_back(*out);
  LC_CHECK_CUDA(cudaMemcpyAsync(*out, s->p, static_cast<size_t>(n) * 4, cudaMemcpyDeviceToDevice, st));
  return 0;
}

// ------------------------------------------------------------------------------------------------ GEMM helpers
struct Ctx {
  lc_denoiser* D;
  cudaStream_t st;
  int run(const GemmArgs& g) const { return D->f32 ? gemm_f32(g, st) : gemm_bf16(g, st); }
  bool fused_qk() const { return !D->f32 && D->fuse_qk; }

  GemmArgs base(const void* A, long long lda, int M, const Lin& L) const {
    GemmArgs g;
    g.A0 = A; g.lda0 = lda; g.W = L.w; g.ldw = L.ldw; g.M = M; g.N = L.out; g.K = L.ldw;
    g.K0 = L.ldw;
    g.epi.bias = L.bias;
    return g;
  }
  // out (storage type T) = act(A W^T + b)
  int lin_T(const void* A, long long lda, int M, const Lin& L, void* out, long long ldo, int act) const {
    GemmArgs g = base(A, lda, M, L);
    g.epi.mode = EPI_STORE; g.epi.act = act; g.epi.out = out; g.epi.ldo = ldo; g.epi.out_f32 = D->f32 ? 1 : 0;
    return run(g);
  }
  int lin_f32(const void* A, long long lda, int M, const Lin& L, float* out, long long ldo, int act) const {
    GemmArgs g = base(A, lda, M, L);
    g.epi.mode = EPI_STORE; g.epi.act = act; g.epi.out = out; g.epi.ldo = ldo; g.epi.out_f32 = 1;
    return run(g);
  }
  // qkv projection into the joint [B, S, 3d] buffer at token offset `tok_off`
  // On the tensor-core path the per-head RMSNorm(q,k) + RoPE of this stream's tokens is fused into the epilogue
  // (seg gives the norm weights and cos/sin tables); the FP32 validation path runs qk_norm_rope afterwards.
  int lin_qkv(const void* A, int M, int rows_per_sample, const Lin& L, void* qkv, int S, int tok_off,
              const RopeSeg* seg) const {
    GemmArgs g = base(A, D->d, M, L);
    g.epi.mode = EPI_STORE; g.epi.out = qkv; g.epi.ldo = 3LL * D->d; g.epi.out_f32 = D->f32 ? 1 : 0;
    g.epi.rows_per_sample = rows_per_sample; g.epi.out_rows_per_sample = S; g.epi.out_row_offset = tok_off;
    if (!D->f32 && seg != nullptr && D->fuse_qk) {
      g.epi.qk_cols = 2 * D->d; g.epi.qk_eps = 1e-7f; g.epi.qk_wq = seg->wq; g.epi.qk_wk = seg->wk;
      g.epi.rope_cs = seg->cos == nullptr ? nullptr
                      : (seg->cos == D->cos_p.as<float>() ? D->cs_p.as<uint32_t>() : D->cs_c.as<uint32_t>());
    }
    return run(g);
  }
  // second row segment of a merged pred+cond launch: rows >= Mp are cond tokens (Nc per sample)
  void cond_segment(EpiParams& ep, int Mp, int Nc, int out_rows_per_sample, int out_row_offset, long long out_base) const {
    ep.seg_rows = Mp; ep.seg_rows_per_sample = Nc; ep.seg_out_rows_per_sample = out_rows_per_sample;
    ep.seg_out_row_offset = out_row_offset; ep.seg_out_base = out_base;
  }
  // single-stream blocks: q|k|v of the pred rows [0, Mp) and the cond rows [Mp, Mp+Mc) of A into the joint qkv buffer
  int lin_qkv_both(const void* A, int Mp, int Np, int Mc, int Nc, const Lin& L, void* qkv, int S) const {
    GemmArgs g = base(A, D->d, Mp + Mc, L);
    g.epi.mode = EPI_STORE; g.epi.out = qkv; g.epi.ldo = 3LL * D->d; g.epi.out_f32 = D->f32 ? 1 : 0;
    g.epi.rows_per_sample = Np; g.epi.out_rows_per_sample = S; g.epi.out_row_offset = 0;
    cond_segment(g.epi, Mp, Nc, S, Np, 0);
    return run(g);
  }
  // single-stream blocks: gated residual update of both streams (rows are identity-mapped, gate per sample)
  int lin_gated_both(const void* A0, long long lda0, int K0, const void* A1, long long lda1, int Mp, int Np, int Mc, int Nc,
                     const Lin& L, float* resid, const float* gate, long long gate_stride) const {
    GemmArgs g = base(A0, lda0, Mp + Mc, L);
    if (A1 != nullptr) { g.A1 = A1; g.lda1 = lda1; g.K0 = K0; }
    g.epi.mode = EPI_GATED_RESID; g.epi.out = resid; g.epi.ldo = D->d; g.epi.rows_per_sample = Np;
    g.epi.out_rows_per_sample = Np; g.epi.gate = gate; g.epi.gate_stride = gate_stride;
    cond_segment(g.epi, Mp, Nc, Nc, 0, Mp);
    return run(g);
  }
  // resid[row] += gate[sample] * (A W^T + b);  A may be two K-segments
  int lin_gated(const void* A0, long long lda0, int K0, const void* A1, long long lda1, int M, int rows_per_sample,
                const Lin& L, float* resid, const float* gate, long long gate_stride) const {
    GemmArgs g = base(A0, lda0, M, L);
    if (A1 != nullptr) { g.A1 = A1; g.lda1 = lda1; g.K0 = K0; }
    g.epi.mode = EPI_GATED_RESID; g.epi.out = resid; g.epi.ldo = D->d; g.epi.rows_per_sample = rows_per_sample;
    g.epi.out_rows_per_sample = rows_per_sample; g.epi.gate = gate; g.epi.gate_stride = gate_stride;
    return run(g);
  }
};

template <typename T>
int attention(lc_denoiser* D, int B, int S, int Np, cudaStream_t st);
template <>
int attention<float>(lc_denoiser* D, int B, int S, int Np, cudaStream_t st) {
  return attention_f32(D->qkv.as<float>(), B, S, D->cfg.num_heads, 128, D->att_p.as<float>(), Np,
                       D->att_p.as<float>() + static_cast<size_t>(B) * D->Np * D->d, st);
}
template <>
int attention<bf16>(lc_denoiser* D, int B, int S, int Np, cudaStream_t st) {
  return attention_bf16(D->qkv.as<bf16>(), B, S, D->cfg.num_heads, 128, D->att_p.as<bf16>(), Np,
                        D->att_p.as<bf16>() + static_cast<size_t>(B) * D->Np * D->d, st);
}

template <typename T>
int prepare_impl(lc_denoiser* D, const float* known, int B, const float* year_emb, int n_ts, cudaStream_t st) {
  Ctx c{D, st};
  const int d = D->d, Nc = D->Nc, Mc = B * Nc;
  // cond tokens -> context_embedder (Conv3d k=1 == per-token linear; embeddings.py:52-59)
  LC_TRY(patchify<T>(known, D->tok_c.as<T>(), B, D->cfg.cond_channels, Nc, D->kp_in, st));
  LC_TRY(c.lin_f32(D->tok_c.p, D->kp_in, Mc, D->c_emb, D->e0.as<float>(), d, ACT_NONE));
  // pooled projection of the refiner + its text embedder (LaDCast_3D_model.py:382-384) are timestep independent
  LC_TRY(token_mean<T>(D->e0.as<float>(), B, Nc, d, D->pooled.as<T>(), st));
  LC_TRY(c.lin_T(D->pooled.p, d, B, D->r_p1, D->tmpA.p, d, ACT_SILU));
  LC_TRY(c.lin_f32(D->tmpA.p, d, B, D->r_p2, D->r_pe.as<float>(), d, ACT_NONE));
  // proj_in(e0)
  LC_TRY(cast_rows<T>(D->e0.as<float>(), D->e0T.as<T>(), static_cast<long long>(Mc) * d, st));
  LC_TRY(c.lin_f32(D->e0T.p, d, Mc, D->r_proj_in, D->e_proj.as<float>(), d, ACT_NONE));
  // date embedding MLP -> (scale, shift) of temb (LaDCast_3D_model.py:958-969)
  D->n_ts = 0;
  if (year_emb != nullptr && D->cfg.incl_time_elapsed) {
    LC_REQUIRE(n_ts == 1 || n_ts == B, "time_elapsed must have 1 or B entries");
    LC_TRY(cast_rows<T>(year_emb, D->yearT.as<T>(), static_cast<long long>(n_ts) * 256, st));
    LC_TRY(c.lin_T(D->yearT.p, 256, n_ts, D->te1, D->tmpB.p, 2 * d, ACT_SILU));
    LC_TRY(c.lin_f32(D->tmpB.p, 2 * d, n_ts, D->te2, D->te_out.as<float>(), 2 * d, ACT_NONE));
    D->n_ts = n_ts;
  }
  D->curB = B;
  return 0;
}

template <typename T>
int forward_impl(lc_denoiser* D, const float* x_in, const float* c_noise, int n_t, float* out, cudaStream_t st) {
  Ctx c{D, st};
  const int d = D->d, B = D->curB, Np = D->Np, Nc = D->Nc, S = D->S;
  const int Mp = B * Np, Mc = B * Nc;
  const int heads = D->cfg.num_heads;
  // the cond-token stream lives directly behind the pred-token stream in every per-token buffer, so that the
  // single-stream blocks (shared weights) can run both streams in one GEMM launch of M = Mp + Mc rows
  float* h = D->h.as<float>();
  float* e = D->e_ptr();
  T* n_p = D->n_p.as<T>();
  T* n_c = n_p + static_cast<size_t>(Mp) * d;
  T* att_c = D->att_p.as<T>() + static_cast<size_t>(Mp) * d;
  T* mlp_c = D->mlp_p.as<T>() + static_cast<size_t>(Mp) * D->cfg.mlp_dim;
  T* qkv = D->qkv.as<T>();
  const long long md = D->mod_dim;
  const float* mod = D->modv.as<float>();

  // ---- embeddings
  LC_TRY(patchify<T>(x_in, D->tok_x.as<T>(), B, D->cfg.in_channels, Np, D->kp_in, st));
  LC_TRY(c.lin_f32(D->tok_x.p, D->kp_in, Mp, D->x_emb, h, d, ACT_NONE));
  LC_CHECK_CUDA(cudaMemcpyAsync(e, D->e_proj.p, static_cast<size_t>(Mc) * d * 4, cudaMemcpyDeviceToDevice, st));
  LC_TRY(timestep_embed<T>(c_noise, n_t, B, D->sincos.as<T>(), st));

  // ---- context refiner (LaDCast_3D_model.py:375-390, 280-302)
  LC_TRY(c.lin_T(D->sincos.p, 256, B, D->r_t1, D->tmpA.p, d, ACT_SILU));
  LC_TRY(c.lin_f32(D->tmpA.p, d, B, D->r_t2, D->r_te.as<float>(), d, ACT_NONE));
  LC_TRY(temb_combine<T>(D->r_te.as<float>(), D->r_pe.as<float>(), nullptr, nullptr, 0, B, d, nullptr,
                         D->r_tembS.as<T>(), st));
  const int n_ref = D->cfg.num_refiner_layers;
  if (n_ref > 0) LC_TRY(c.lin_f32(D->r_tembS.p, d, B, D->r_gates, D->gates.as<float>(), 2LL * d * n_ref, ACT_NONE));
  for (int i = 0; i < n_ref; ++i) {
    const RefinerW& w = D->refiner[i];
    const float* g_msa = D->gates.as<float>() + 2LL * d * i;
    const float* g_mlp = g_msa + d;
    const long long gs = 2LL * d * n_ref;
    LC_TRY(layernorm_modulate<T>(e, n_c, Mc, d, 1e-7f, Nc, nullptr, nullptr, 0, w.n1w, w.n1b, st));
    RopeSeg seg;
    seg.start = 0; seg.len = Nc; seg.wq = w.nq; seg.wk = w.nk; seg.cos = D->cos_c.as<float>(); seg.sin = D->sin_c.as<float>();
    seg.cs = D->cs_c.as<uint32_t>();
    LC_TRY(c.lin_qkv(n_c, Mc, Nc, w.qkv, qkv, Nc, 0, &seg));
    if (!c.fused_qk()) LC_TRY(qk_norm_rope<T>(qkv, 3LL * d, B, Nc, heads, 128, 1e-7f, &seg, 1, st));
    LC_TRY(attention<T>(D, B, Nc, 0, st));  // all tokens -> att_c
    LC_TRY(gated_add<T>(e, att_c, g_msa, gs, Mc, d, Nc, st));
    LC_TRY(layernorm_modulate<T>(e, n_c, Mc, d, 1e-7f, Nc, nullptr, nullptr, 0, w.n2w, w.n2b, st));
    LC_TRY(c.lin_T(n_c, d, Mc, w.ff0, mlp_c, w.ff0.out, ACT_SILU));
    LC_TRY(c.lin_gated(mlp_c, w.ff0.out, 0, nullptr, 0, Mc, Nc, w.ff2, e, g_mlp, gs));
  }

  // ---- temb = time_text_embed(t, mean(e)) [* (1+scale_date) + shift_date]  (:953-969)
  LC_TRY(c.lin_T(D->sincos.p, 256, B, D->t1, D->tmpA.p, d, ACT_SILU));
  LC_TRY(c.lin_f32(D->tmpA.p, d, B, D->t2, D->t_te.as<float>(), d, ACT_NONE));
  LC_TRY(token_mean<T>(e, B, Nc, d, D->pooled.as<T>(), st));
  LC_TRY(c.lin_T(D->pooled.p, d, B, D->p1, D->tmpA.p, d, ACT_SILU));
  LC_TRY(c.lin_f32(D->tmpA.p, d, B, D->p2, D->pe.as<float>(), d, ACT_NONE));
  {
    const float* sc = D->n_ts ? D->te_out.as<float>() : nullptr;
    const float* sh = D->n_ts ? D->te_out.as<float>() + d : nullptr;
    LC_TRY(temb_combine<T>(D->t_te.as<float>(), D->pe.as<float>(), sc, sh, D->n_ts == 1 ? 0 : 2LL * d, B, d,
                           D->temb.as<float>(), D->tembS.as<T>(), st));
  }
  // ---- every AdaLN modulation vector of the call in one GEMM
  LC_TRY(c.lin_f32(D->tembS.p, d, B, D->mod, D->modv.as<float>(), md, ACT_NONE));

  // ---- dual-stream blocks (:514-566)
  for (int i = 0; i < D->cfg.num_layers; ++i) {
    const DualW& w = D->dual[i];
    const float* mh = mod + 12LL * d * i;  // shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp
    const float* mc = mh + 6LL * d;
    LC_TRY(layernorm_modulate<T>(h, n_p, Mp, d, 1e-6f, Np, mh + d, mh, md, nullptr, nullptr, st));
    LC_TRY(layernorm_modulate<T>(e, n_c, Mc, d, 1e-6f, Nc, mc + d, mc, md, nullptr, nullptr, st));
    RopeSeg segs[2];
    segs[0].start = 0; segs[0].len = Np; segs[0].wq = w.nq; segs[0].wk = w.nk;
    segs[0].cos = D->cos_p.as<float>(); segs[0].sin = D->sin_p.as<float>(); segs[0].cs = D->cs_p.as<uint32_t>();
    segs[1].start = Np; segs[1].len = Nc; segs[1].wq = w.naq; segs[1].wk = w.nak;  // no RoPE on cond (quirk C-3)
    LC_TRY(c.lin_qkv(n_p, Mp, Np, w.qkv, qkv, S, 0, &segs[0]));
    LC_TRY(c.lin_qkv(n_c, Mc, Nc, w.add_qkv, qkv, S, Np, &segs[1]));
    if (!c.fused_qk()) LC_TRY(qk_norm_rope<T>(qkv, 3LL * d, B, S, heads, 128, 1e-7f, segs, 2, st));
    LC_TRY(attention<T>(D, B, S, Np, st));
    LC_TRY(c.lin_gated(D->att_p.p, d, 0, nullptr, 0, Mp, Np, w.to_out, h, mh + 2 * d, md));
    LC_TRY(c.lin_gated(att_c, d, 0, nullptr, 0, Mc, Nc, w.to_add_out, e, mc + 2 * d, md));
    LC_TRY(layernorm_modulate<T>(h, n_p, Mp, d, 1e-7f, Np, mh + 4 * d, mh + 3 * d, md, nullptr, nullptr, st));
    LC_TRY(layernorm_modulate<T>(e, n_c, Mc, d, 1e-7f, Nc, mc + 4 * d, mc + 3 * d, md, nullptr, nullptr, st));
    LC_TRY(c.lin_T(n_p, d, Mp, w.ff0, D->mlp_p.p, w.ff0.out, ACT_GELU_TANH));
    LC_TRY(c.lin_T(n_c, d, Mc, w.ffc0, mlp_c, w.ffc0.out, ACT_GELU_TANH));
    LC_TRY(c.lin_gated(D->mlp_p.p, w.ff0.out, 0, nullptr, 0, Mp, Np, w.ff2, h, mh + 5 * d, md));
    LC_TRY(c.lin_gated(mlp_c, w.ffc0.out, 0, nullptr, 0, Mc, Nc, w.ffc2, e, mc + 5 * d, md));
  }

  // ---- single-stream blocks (:426-468): both streams share weights and modulation
  const float* ms_base = mod + 12LL * d * D->cfg.num_layers;
  for (int i = 0; i < D->cfg.num_single_layers; ++i) {
    const SingleW& w = D->single[i];
    const float* ms = ms_base + 3LL * d * i;  // shift, scale, gate
    const bool merged = D->merge_streams;
    if (merged) {  // same modulation for both streams: one pass over the Mp + Mc rows of h | e
      LC_TRY(layernorm_modulate<T>(h, n_p, Mp + Mc, d, 1e-6f, Np, ms + d, ms, md, nullptr, nullptr, st, Mp, Nc));
    } else {
      LC_TRY(layernorm_modulate<T>(h, n_p, Mp, d, 1e-6f, Np, ms + d, ms, md, nullptr, nullptr, st));
      LC_TRY(layernorm_modulate<T>(e, n_c, Mc, d, 1e-6f, Nc, ms + d, ms, md, nullptr, nullptr, st));
    }
    RopeSeg segs[2];
    segs[0].start = 0; segs[0].len = Np; segs[0].wq = w.nq; segs[0].wk = w.nk;
    segs[0].cos = D->cos_p.as<float>(); segs[0].sin = D->sin_p.as<float>(); segs[0].cs = D->cs_p.as<uint32_t>();
    segs[1].start = Np; segs[1].len = Nc; segs[1].wq = w.nq; segs[1].wk = w.nk;
    segs[1].cos = D->cos_c.as<float>(); segs[1].sin = D->sin_c.as<float>(); segs[1].cs = D->cs_c.as<uint32_t>();
    if (c.fused_qk() || !merged) {  // (fused: per-stream rotation tables live in the epilogue)
      LC_TRY(c.lin_qkv(n_p, Mp, Np, w.qkv, qkv, S, 0, &segs[0]));
      LC_TRY(c.lin_qkv(n_c, Mc, Nc, w.qkv, qkv, S, Np, &segs[1]));
    } else {
      LC_TRY(c.lin_qkv_both(n_p, Mp, Np, Mc, Nc, w.qkv, qkv, S));
    }
    // both streams share the weights (:426-468): one launch over the Mp + Mc rows of n_p | n_c
    if (merged) {
      LC_TRY(c.lin_T(n_p, d, Mp + Mc, w.mlp, D->mlp_p.p, w.mlp.out, ACT_GELU_TANH));
    } else {
      LC_TRY(c.lin_T(n_p, d, Mp, w.mlp, D->mlp_p.p, w.mlp.out, ACT_GELU_TANH));
      LC_TRY(c.lin_T(n_c, d, Mc, w.mlp, mlp_c, w.mlp.out, ACT_GELU_TANH));
    }
    if (!c.fused_qk()) LC_TRY(qk_norm_rope<T>(qkv, 3LL * d, B, S, heads, 128, 1e-7f, segs, 2, st));
    LC_TRY(attention<T>(D, B, S, Np, st));
    // proj_out over [attn | mlp] (K = d + mlp_dim) read from two buffers, gate, + residual
    if (merged) {
      LC_TRY(c.lin_gated_both(D->att_p.p, d, d, D->mlp_p.p, w.mlp.out, Mp, Np, Mc, Nc, w.proj_out, h, ms + 2 * d, md));
    } else {
      LC_TRY(c.lin_gated(D->att_p.p, d, d, D->mlp_p.p, w.mlp.out, Mp, Np, w.proj_out, h, ms + 2 * d, md));
      LC_TRY(c.lin_gated(att_c, d, d, mlp_c, w.mlp.out, Mc, Nc, w.proj_out, e, ms + 2 * d, md));
    }
  }

  // ---- norm_out (AdaLayerNormContinuous: chunk order scale, shift) + proj_out + unpatchify (:1044-1062)
  const float* mo = ms_base + 3LL * d * D->cfg.num_single_layers;
  LC_TRY(layernorm_modulate<T>(h, n_p, Mp, d, 1e-7f, Np, mo, mo + d, md, nullptr, nullptr, st));
  {
    GemmArgs g = c.base(n_p, d, Mp, D->proj_out);
    g.epi.mode = EPI_UNPATCHIFY; g.epi.out = out; g.epi.rows_per_sample = Np; g.epi.n_valid = D->cfg.out_channels;
    LC_TRY(c.run(g));
  }
  return 0;
}

int finalize_impl(lc_denoiser* D, cudaStream_t st) {
  const int d = D->d, mlp = D->cfg.mlp_dim;
  auto L = [&](const std::vector<std::string>& names, int in, int ldw, Lin* out) { return make_lin(D, names, in, ldw, true, out, st); };
  LC_TRY(L({"x_embedder.proj"}, D->cfg.in_channels, D->kp_in, &D->x_emb));
  LC_TRY(L({"context_embedder.proj"}, D->cfg.cond_channels, D->kp_in, &D->c_emb));
  LC_TRY(L({"context_refiner.time_text_embed.timestep_embedder.linear_1"}, 256, 256, &D->r_t1));
  LC_TRY(L({"context_refiner.time_text_embed.timestep_embedder.linear_2"}, d, d, &D->r_t2));
  LC_TRY(L({"context_refiner.time_text_embed.text_embedder.linear_1"}, d, d, &D->r_p1));
  LC_TRY(L({"context_refiner.time_text_embed.text_embedder.linear_2"}, d, d, &D->r_p2));
  LC_TRY(L({"context_refiner.proj_in"}, d, d, &D->r_proj_in));
  LC_TRY(L({"time_text_embed.timestep_embedder.linear_1"}, 256, 256, &D->t1));
  LC_TRY(L({"time_text_embed.timestep_embedder.linear_2"}, d, d, &D->t2));
  LC_TRY(L({"time_text_embed.text_embedder.linear_1"}, d, d, &D->p1));
  LC_TRY(L({"time_text_embed.text_embedder.linear_2"}, d, d, &D->p2));
  if (D->cfg.incl_time_elapsed) {
    LC_TRY(L({"time_elapsed_embed.linear_1"}, 256, 256, &D->te1));
    LC_TRY(L({"time_elapsed_embed.linear_2"}, 2 * d, 2 * d, &D->te2));
  }
  std::vector<std::string> gate_names, mod_names;
  D->refiner.resize(D->cfg.num_refiner_layers);
  for (int i = 0; i < D->cfg.num_refiner_layers; ++i) {
    const std::string p = "context_refiner.token_refiner.refiner_blocks." + std::to_string(i);
    RefinerW& w = D->refiner[i];
    LC_TRY(vec(D, p + ".norm1.weight", d, &w.n1w, st));
    LC_TRY(vec(D, p + ".norm1.bias", d, &w.n1b, st));
    LC_TRY(vec(D, p + ".norm2.weight", d, &w.n2w, st));
    LC_TRY(vec(D, p + ".norm2.bias", d, &w.n2b, st));
    LC_TRY(vec(D, p + ".attn.norm_q.weight", 128, &w.nq, st));
    LC_TRY(vec(D, p + ".attn.norm_k.weight", 128, &w.nk, st));
    LC_TRY(L({p + ".attn.to_q", p + ".attn.to_k", p + ".attn.to_v"}, d, d, &w.qkv));
    LC_TRY(L({p + ".ff.net.0.proj"}, d, d, &w.ff0));
    LC_TRY(L({p + ".ff.net.2"}, mlp, mlp, &w.ff2));
    gate_names.push_back(p + ".norm_out.linear");
  }
  if (!gate_names.empty()) LC_TRY(L(gate_names, d, d, &D->r_gates));
  D->dual.resize(D->cfg.num_layers);
  for (int i = 0; i < D->cfg.num_layers; ++i) {
    const std::string p = "transformer_blocks." + std::to_string(i);
    DualW& w = D->dual[i];
    LC_TRY(L({p + ".attn.to_q", p + ".attn.to_k", p + ".attn.to_v"}, d, d, &w.qkv));
    LC_TRY(L({p + ".attn.add_q_proj", p + ".attn.add_k_proj", p + ".attn.add_v_proj"}, d, d, &w.add_qkv));
    LC_TRY(L({p + ".attn.to_out.0"}, d, d, &w.to_out));
    LC_TRY(L({p + ".attn.to_add_out"}, d, d, &w.to_add_out));
    LC_TRY(L({p + ".ff.net.0.proj"}, d, d, &w.ff0));
    LC_TRY(L({p + ".ff.net.2"}, mlp, mlp, &w.ff2));
    LC_TRY(L({p + ".ff_context.net.0.proj"}, d, d, &w.ffc0));
    LC_TRY(L({p + ".ff_context.net.2"}, mlp, mlp, &w.ffc2));
    LC_TRY(vec(D, p + ".attn.norm_q.weight", 128, &w.nq, st));
    LC_TRY(vec(D, p + ".attn.norm_k.weight", 128, &w.nk, st));
    LC_TRY(vec(D, p + ".attn.norm_added_q.weight", 128, &w.naq, st));
    LC_TRY(vec(D, p + ".attn.norm_added_k.weight", 128, &w.nak, st));
    mod_names.push_back(p + ".norm1.linear");
    mod_names.push_back(p + ".norm1_context.linear");
  }
  D->single.resize(D->cfg.num_single_layers);
  for (int i = 0; i < D->cfg.num_single_layers; ++i) {
    const std::string p = "single_transformer_blocks." + std::to_string(i);
    SingleW& w = D->single[i];
    LC_TRY(L({p + ".attn.to_q", p + ".attn.to_k", p + ".attn.to_v"}, d, d, &w.qkv));
    LC_TRY(L({p + ".proj_mlp"}, d, d, &w.mlp));
    LC_TRY(L({p + ".proj_out"}, d + mlp, d + mlp, &w.proj_out));
    LC_TRY(vec(D, p + ".attn.norm_q.weight", 128, &w.nq, st));
    LC_TRY(vec(D, p + ".attn.norm_k.weight", 128, &w.nk, st));
    mod_names.push_back(p + ".norm.linear");
  }
  mod_names.push_back("norm_out.linear");
  LC_TRY(L(mod_names, d, d, &D->mod));
  D->mod_dim = D->mod.out;
  LC_REQUIRE(D->mod_dim == 12 * d * D->cfg.num_layers + 3 * d * D->cfg.num_single_layers + 2 * d, "modulation size");
  LC_TRY(L({"proj_out"}, d, d, &D->proj_out));
  LC_CHECK_CUDA(cudaStreamSynchronize(st));
  for (auto& kv : D->staged) cudaFree(kv.second.p);
  D->staged.clear();
  D->finalized = true;
  return 0;
}

}  // namespace
}  // namespace lc

// ================================================================================================ C ABI
extern "C" {

int lc_denoiser_create(const lc_denoiser_cfg* cfg, lc_denoiser** out) {
  LC_REQUIRE(cfg != nullptr && out != nullptr, "null argument");
  LC_REQUIRE(cfg->head_dim == 128, "attention_head_dim must be 128 (sum(rope_axes_dim))");
  LC_REQUIRE(cfg->in_channels <= 96 && cfg->cond_channels <= 96 && cfg->out_channels <= 128, "channel count too large");
  LC_REQUIRE(cfg->precision == LC_PRECISION_BF16 || cfg->precision == LC_PRECISION_F32, "unknown precision");
  LC_REQUIRE(cfg->mlp_dim % 64 == 0, "mlp_dim must be a multiple of 64");
  lc_denoiser* D = new lc_denoiser();
  D->cfg = *cfg;
  D->f32 = cfg->precision == LC_PRECISION_F32;
  D->esz = D->f32 ? 4 : 2;
  D->d = cfg->num_heads * cfg->head_dim;
  const char* fq = getenv("LADCAST_B200_FUSE_QK");
  D->fuse_qk = fq != nullptr && fq[0] == '1';
  const char* mg = getenv("LADCAST_B200_MERGE_STREAMS");
  D->merge_streams = !(mg != nullptr && mg[0] == '0');
  *out = D;
  return 0;
}

void lc_denoiser_destroy(lc_denoiser* D) {
  if (!D) return;
  for (void* p : D->owned) cudaFree(p);
  for (auto& kv : D->staged) cudaFree(kv.second.p);
  DevBuf* bufs[] = {&D->cos_p, &D->sin_p, &D->cos_c, &D->sin_c, &D->cs_p, &D->cs_c, &D->tok_x, &D->tok_c, &D->h, &D->e0, &D->e0T,
                    &D->e_proj, &D->n_p, &D->qkv, &D->att_p, &D->mlp_p, &D->sincos,
                    &D->tmpA, &D->tmpB, &D->r_te, &D->r_pe, &D->r_tembS, &D->gates, &D->t_te, &D->pooled, &D->pe,
                    &D->temb, &D->tembS, &D->modv, &D->te_out, &D->yearT};
  for (DevBuf* b : bufs) b->release();
  delete D;
}

int lc_denoiser_load(lc_denoiser* D, const char* key, const float* data, const int64_t* shape, int ndim, void* stream) {
  LC_REQUIRE(D && key && data && shape, "null argument");
  LC_REQUIRE(!D->finalized, "lc_denoiser_load after finalize");
  Staged s;
  s.numel = 1;
  for (int i = 0; i < ndim; ++i) {
    s.shape.push_back(shape[i]);
    s.numel *= shape[i];
  }
  LC_CHECK_CUDA(cudaMalloc(reinterpret_cast<void**>(&s.p), static_cast<size_t>(s.numel) * 4));
  LC_CHECK_CUDA(cudaMemcpyAsync(s.p, data, static_cast<size_t>(s.numel) * 4, cudaMemcpyDeviceToDevice,
                                static_cast<cudaStream_t>(stream)));
  auto it = D->staged.find(key);
  if (it != D->staged.end()) cudaFree(it->second.p);
  D->staged[key] = s;
  return 0;
}

int lc_denoiser_finalize(lc_denoiser* D, void* stream) {
  LC_REQUIRE(D && !D->finalized, "finalize called twice or on null handle");
  return finalize_impl(D, static_cast<cudaStream_t>(stream));
}

int lc_denoiser_set_geometry(lc_denoiser* D, int max_batch, int t_in, int t_out, int height, int width,
                             const float* cos_pred, const float* sin_pred, const float* cos_cond,
                             const float* sin_cond, void* stream) {
  LC_REQUIRE(D && D->finalized, "set_geometry before finalize");
  LC_REQUIRE(max_batch > 0 && t_in > 0 && t_out > 0 && height > 0 && width > 0, "bad geometry");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  D->maxB = max_batch; D->T_in = t_in; D->T_out = t_out; D->H = height; D->W = width;
  D->Np = t_out * height * width; D->Nc = t_in * height * width; D->S = D->Np + D->Nc;
  const size_t B = max_batch, d = D->d, e = D->esz, Np = D->Np, Nc = D->Nc, S = D->S, mlp = D->cfg.mlp_dim;
  LC_TRY(D->cos_p.alloc(Np * 128 * 4)); LC_TRY(D->sin_p.alloc(Np * 128 * 4));
  LC_TRY(D->cos_c.alloc(Nc * 128 * 4)); LC_TRY(D->sin_c.alloc(Nc * 128 * 4));
  LC_CHECK_CUDA(cudaMemcpyAsync(D->cos_p.p, cos_pred, Np * 128 * 4, cudaMemcpyDeviceToDevice, st));
  LC_CHECK_CUDA(cudaMemcpyAsync(D->sin_p.p, sin_pred, Np * 128 * 4, cudaMemcpyDeviceToDevice, st));
  LC_CHECK_CUDA(cudaMemcpyAsync(D->cos_c.p, cos_cond, Nc * 128 * 4, cudaMemcpyDeviceToDevice, st));
  LC_CHECK_CUDA(cudaMemcpyAsync(D->sin_c.p, sin_cond, Nc * 128 * 4, cudaMemcpyDeviceToDevice, st));
  LC_TRY(D->cs_p.alloc(Np * 64 * 4)); LC_TRY(D->cs_c.alloc(Nc * 64 * 4));
  LC_TRY(pack_rope_pairs(D->cos_p.as<float>(), D->sin_p.as<float>(), D->cs_p.as<uint32_t>(), static_cast<int>(Np), st));
  LC_TRY(pack_rope_pairs(D->cos_c.as<float>(), D->sin_c.as<float>(), D->cs_c.as<uint32_t>(), static_cast<int>(Nc), st));
  LC_TRY(D->tok_x.alloc(B * Np * D->kp_in * e)); LC_TRY(D->tok_c.alloc(B * Nc * D->kp_in * e));
  LC_TRY(D->h.alloc(B * S * d * 4));
  LC_TRY(D->e0.alloc(B * Nc * d * 4)); LC_TRY(D->e0T.alloc(B * Nc * d * e)); LC_TRY(D->e_proj.alloc(B * Nc * d * 4));
  LC_TRY(D->n_p.alloc(B * S * d * e));
  LC_TRY(D->qkv.alloc(B * S * 3 * d * e));
  LC_TRY(D->att_p.alloc(B * S * d * e));
  LC_TRY(D->mlp_p.alloc(B * S * mlp * e));
  LC_TRY(D->sincos.alloc(B * 256 * e)); LC_TRY(D->tmpA.alloc(B * d * e)); LC_TRY(D->tmpB.alloc(B * 2 * d * e));
  LC_TRY(D->r_te.alloc(B * d * 4)); LC_TRY(D->r_pe.alloc(B * d * 4)); LC_TRY(D->r_tembS.alloc(B * d * e));
  LC_TRY(D->gates.alloc(B * 2 * d * (D->cfg.num_refiner_layers > 0 ? D->cfg.num_refiner_layers : 1) * 4));
  LC_TRY(D->t_te.alloc(B * d * 4)); LC_TRY(D->pooled.alloc(B * d * e)); LC_TRY(D->pe.alloc(B * d * 4));
  LC_TRY(D->temb.alloc(B * d * 4)); LC_TRY(D->tembS.alloc(B * d * e));
  LC_TRY(D->modv.alloc(B * static_cast<size_t>(D->mod_dim) * 4));
  LC_TRY(D->te_out.alloc(B * 2 * d * 4)); LC_TRY(D->yearT.alloc(B * 256 * e));
  D->curB = 0;
  return 0;
}

int lc_denoiser_prepare(lc_denoiser* D, const float* known, int batch, const float* year_emb, int n_ts, void* stream) {
  LC_REQUIRE(D && D->maxB > 0, "prepare before set_geometry");
  LC_REQUIRE(batch > 0 && batch <= D->maxB, "batch exceeds max_batch given to set_geometry");
  LC_REQUIRE(known != nullptr, "known_latents must be provided");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return D->f32 ? prepare_impl<float>(D, known, batch, year_emb, n_ts, st)
                : prepare_impl<bf16>(D, known, batch, year_emb, n_ts, st);
}

int lc_denoiser_forward(lc_denoiser* D, const float* x_in, const float* c_noise, int n_t, float* out, void* stream) {
  LC_REQUIRE(D && D->curB > 0, "forward before prepare");
  LC_REQUIRE(x_in && c_noise && out && x_in != out, "bad forward arguments");
  LC_REQUIRE(n_t == 1 || n_t == D->curB, "timestep must have 1 or B entries");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return D->f32 ? forward_impl<float>(D, x_in, c_noise, n_t, out, st) : forward_impl<bf16>(D, x_in, c_noise, n_t, out, st);
}

int lc_denoiser_debug_read(lc_denoiser* D, const char* name, float* out, int64_t max_elems, void* stream) {
  LC_REQUIRE(D && name && out, "null argument");
  const std::string n(name);
  const float* src = nullptr;
  int64_t cnt = 0;
  if (n == "h") { src = D->h.as<float>(); cnt = static_cast<int64_t>(D->curB) * D->Np * D->d; }
  else if (n == "e") { src = D->e_ptr(); cnt = static_cast<int64_t>(D->curB) * D->Nc * D->d; }
  else if (n == "temb") { src = D->temb.as<float>(); cnt = static_cast<int64_t>(D->curB) * D->d; }
  else if (n == "mod") { src = D->modv.as<float>(); cnt = static_cast<int64_t>(D->curB) * D->mod_dim; }
  LC_REQUIRE(src != nullptr, "unknown debug buffer '" + n + "'");
  if (cnt > max_elems) cnt = max_elems;
  LC_CHECK_CUDA(cudaMemcpyAsync(out, src, static_cast<size_t>(cnt) * 4, cudaMemcpyDeviceToDevice,
                                static_cast<cudaStream_t>(stream)));
  return 0;
}

}  // extern "C"
