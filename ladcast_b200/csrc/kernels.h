// Internal launch functions of the non-GEMM kernels.  T = activation storage type (bf16 in the production
// mode, float in the FP32 validation mode); all statistics/accumulation are fp32.
#pragma once
#include "common.cuh"

namespace lc {

struct RopeSeg {
  int start = 0, len = 0;         // token range [start, start+len) inside a sample's sequence
  const float* wq = nullptr;      // RMSNorm weights [head_dim]
  const float* wk = nullptr;
  const float* cos = nullptr;     // [len, head_dim] or null (no rotation: dual-stream cond tokens)
  const float* sin = nullptr;
  const uint32_t* cs = nullptr;   // optional packed [len, 64] half2 (cos, sin) per rotation pair (bf16 kernel)
};

template <typename T>
int layernorm_modulate(const float* x, T* out, int M, int d, float eps, int rows_per_sample, const float* scale,
                       const float* shift, long long mod_stride, const float* w, const float* b, cudaStream_t s, int seg_rows = 0,
                       int seg_rows_per_sample = 1);
int pack_rope_pairs(const float* cos, const float* sin, uint32_t* out, int n_tokens, cudaStream_t s);
template <typename T>
int qk_norm_rope(T* qkv, long long ld, int B, int S, int heads, int head_dim, float eps, const RopeSeg* segs, int nseg,
                 cudaStream_t s);
template <typename T>
int patchify(const float* x, T* out, int B, int C, int THW, int Kp, cudaStream_t s);
template <typename T>
int timestep_embed(const float* t, int n_t, int B, T* out, cudaStream_t s);
template <typename T>
int token_mean(const float* x, int B, int N, int d, T* out, cudaStream_t s);
template <typename T>
int gated_add(float* h, const T* a, const float* gate, long long gate_stride, int M, int d, int rows_per_sample,
              cudaStream_t s);
// out = (a + b) * (1 + sc) + sh (sc/sh optional, row stride sc_stride, 0 = broadcast); writes f32 and/or silu() as T
template <typename T>
int temb_combine(const float* a, const float* b, const float* sc, const float* sh, long long sc_stride, int B, int d,
                 float* out_f32, T* out_silu, cudaStream_t s);
template <typename T>
int cast_rows(const float* x, T* out, long long n, cudaStream_t s);

// attention over a joint sequence stored token-major as [B, S, 3*d] (q | k | v); the output is written
// token-major and split: tokens [0, Np) -> out_p [B*Np, d], tokens [Np, S) -> out_c [B*(S-Np), d].
int attention_f32(const float* qkv, int B, int S, int heads, int head_dim, float* out_p, int Np, float* out_c,
                  cudaStream_t s);
int attention_set_trace(long long* buf);  // tuning aid: timeline of CTA 0 (nullptr = off)
int attention_bf16(const bf16* qkv, int B, int S, int heads, int head_dim, bf16* out_p, int Np, bf16* out_c,
                   cudaStream_t s);

// scheduler (diffusers EDMDPMSolverMultistepScheduler.step + scale_model_input of the NEXT step, fused)
struct SchedCoef {
  float c_skip, c_out;     // x0 = c_skip*x + c_out*F
  float a_x, a_x0, a_d;    // x' = a_x*x + a_x0*x0 + a_d*(x0 - x0_prev)   (a_d = 0 for first-order steps)
  float c_in_next;         // x_in' = x' * c_in_next (0 -> not written)
};
int sched_dpmpp2m_step(const float* f, float* x, float* x0_prev, float* x_in_next, long long n, SchedCoef c,
                       cudaStream_t s);
int sched_heun_step(const float* f, double* x, double* x_hat, double* d_cur, float* x_in_next, long long n, int phase,
                    double t_cur, double t_next, double c_skip, double c_out, double c_in_next, cudaStream_t s);

// x_in = x * c_in (scale_model_input of the first step); Heun prologue x = f64(noise) * t0, x_in = f32(x * c_in)
int sched_scale_input(const float* x, float* x_in, long long n, float c_in, cudaStream_t s);
int sched_heun_init(const float* noise, double* x, float* x_in, long long n, double t0, double c_in, cudaStream_t s);
int sched_heun_churn(double* x, const double* noise, float* x_in, long long n, double k, double c_in, cudaStream_t s);
// AR feedback: next conditioning frames + optional de-normalised copy of a sampler output [B, C, T, hw]
int latent_feedback(const float* samples, float* known, float* phys, const float* mean, const float* stdv, float target,
                    int B, int C, int T, int t_in, int hw, cudaStream_t s);

}  // namespace lc
