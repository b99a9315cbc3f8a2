/* ladcast_b200 — C ABI of the B200-native (sm_100a) LaDCast ensemble-rollout hot path.
 *
 * The reference (tonyzyl/ladcast) is pure Python/PyTorch and has no FFI boundary; its boundary for this path is the
 * set of Python call signatures listed next to each entry point below (paths relative to the reference's
 * `ladcast/` package).  The Python drop-ins in `ladcast_b200/` bind these symbols with ctypes
 * (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *   - every function returns 0 on success, a negative code on failure; lc_last_error() returns a thread-local
 *     human-readable message for the last failure on the calling thread;
 *   - all tensor arguments are DEVICE pointers unless the name ends in `_host`; the caller owns every I/O buffer;
 *     the library owns weights and workspace;
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered, no host synchronisation and no
 *     allocation happens inside *_forward / *_step / *_decode / *_accumulate calls;
 *   - a handle is not thread-safe; distinct handles are independent.
 */
#ifndef LADCAST_B200_H_
#define LADCAST_B200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define LC_API __attribute__((visibility("default")))
#else
#define LC_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define LC_PRECISION_BF16 0 /* tcgen05 bf16 x bf16 -> f32 tensor-core path (production) */
#define LC_PRECISION_F32 1  /* SIMT fp32 validation path (rel-L2 <= 1e-4 vs the reference) */

LC_API int lc_version(void);
LC_API const char* lc_last_error(void);
/* number of CUDA kernels this library has launched in this process (bench.py's gpu_launches) */
LC_API long long lc_launch_count(void);
/* bench instrumentation: bracket every tensor-core launch (class 0 GEMM, 1 attention, 2 sphere-conv) with CUDA
 * events on its stream; lc_prof_collect sums per-class milliseconds, algorithmic FLOPs and launch counts (arrays of
 * 3) and synchronises on the recorded events. */
LC_API int lc_prof_enable(int on);
LC_API int lc_prof_collect(double* ms, double* flops, long long* launches);
/* the same for EVERY kernel class of the library (lc_prof_num_classes() entries per array; names from
 * lc_prof_class_name): milliseconds, algorithmic FLOPs, algorithmic BYTES (HBM-bound kernels) and launch counts */
LC_API int lc_prof_num_classes(void);
LC_API const char* lc_prof_class_name(int cls);
LC_API int lc_prof_collect_all(double* ms, double* flops, double* bytes, long long* launches);
/* tuning aids (tools/gemm_trace.py, tools/attn_trace.py): device buffer of clock64 stamps written by CTA (pair) 0 of
 * the tcgen05 GEMM / attention kernels; NULL switches the trace off.  Not part of the product path. */
LC_API int lc_debug_gemm_trace(void* device_buf);
LC_API int lc_debug_attention_trace(void* device_buf);

/* ------------------------------------------------------------------------------------------------------------
 * Denoiser — replaces LaDCastTransformer3DModel.__init__/forward (models/LaDCast_3D_model.py:624-650, 833-1071)
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct lc_denoiser lc_denoiser;

typedef struct {
  int in_channels;        /* 84 */
  int out_channels;       /* 84 */
  int cond_channels;      /* conditioning_tensor_in_channels, 84 */
  int num_heads;          /* 12 (375M) / 16 (1.6B) */
  int head_dim;           /* 128 */
  int num_layers;         /* dual-stream blocks   */
  int num_single_layers;  /* single-stream blocks */
  int num_refiner_layers; /* context refiner blocks */
  int mlp_dim;            /* int(hidden * mlp_ratio) */
  int incl_time_elapsed;  /* 1 if time_elapsed_embed exists */
  int precision;          /* LC_PRECISION_* */
} lc_denoiser_cfg;

LC_API int lc_denoiser_create(const lc_denoiser_cfg* cfg, lc_denoiser** out);
LC_API void lc_denoiser_destroy(lc_denoiser* h);

/* Checkpoint loading: one call per state-dict entry of the V0.1.X diffusers-format checkpoint (fp32 tensors,
 * key names as in the reference's state_dict), then finalize() re-packs them into fused,
 * padded library-owned buffers (q|k|v fused, all AdaLN linears fused into one modulation matrix, ...). */
LC_API int lc_denoiser_load(lc_denoiser* h, const char* key, const float* data, const int64_t* shape, int ndim, void* stream);
LC_API int lc_denoiser_finalize(lc_denoiser* h, void* stream);

/* Geometry + RoPE tables (LaDCastRotaryPosEmbed_from_grid, models/embeddings.py:252-327; forward :885-938).
 * cos/sin tables are [T*H*W, head_dim] fp32 for pred (T_out) and cond (T_in) tokens.  Allocates the workspace
 * for up to max_batch members (the only allocating call besides create/load/finalize). */
LC_API int lc_denoiser_set_geometry(lc_denoiser* h, int max_batch, int t_in, int t_out, int height, int width,
                             const float* cos_pred, const float* sin_pred, const float* cos_cond,
                             const float* sin_cond, void* stream);

/* Step-invariant work of one AR step (hoisted out of the num_inference_steps loop): context_embedder(known),
 * its token mean, refiner proj_in, the refiner's text embedder, and the date MLP of
 * get_year_sincos_embedding -> time_elapsed_embed (forward :943-969).  known: [B, C, T_in, H, W] fp32;
 * year_emb: [n_ts, 256] fp32 (n_ts = 1 broadcast or B) or NULL when time_elapsed is None. */
LC_API int lc_denoiser_prepare(lc_denoiser* h, const float* known, int batch, const float* year_emb, int n_ts, void* stream);

/* One denoiser evaluation F(x_in, c_noise | known, date).  x_in/out: [B, C, T_out, H, W] fp32; c_noise: [n_t]
 * fp32 with n_t = B (pipeline_AR.py:92) or 1 (edm_sampler.py:87, broadcast).  out must not alias x_in. */
LC_API int lc_denoiser_forward(lc_denoiser* h, const float* x_in, const float* c_noise, int n_t, float* out, void* stream);

/* Debug tap: copies an internal buffer after a forward ("h", "e", "temb", "mod") to `out` as fp32. */
LC_API int lc_denoiser_debug_read(lc_denoiser* h, const char* name, float* out, int64_t max_elems, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Scheduler — replaces diffusers.EDMDPMSolverMultistepScheduler.step (+ scale_model_input of the next step)
 * as called at pipelines/pipeline_AR.py:87-102, and the Heun update of pipelines/edm_sampler.py:65-113.
 * ---------------------------------------------------------------------------------------------------------- */
/* x0 = c_skip*x + c_out*f ; x <- a_x*x + a_x0*x0 + a_d*(x0 - x0_prev) ; x0_prev <- x0 ;
 * x_in_next <- x*c_in_next (skipped when x_in_next is NULL).  n elements, n % 4 == 0. */
LC_API int lc_sched_dpmpp2m_step(const float* f, float* x, float* x0_prev, float* x_in_next, int64_t n, float c_skip,
                          float c_out, float a_x, float a_x0, float a_d, float c_in_next, void* stream);
/* x_in = x * c_in: scale_model_input of the FIRST step (pipeline_AR.py:90); later steps get it from the fused step */
LC_API int lc_sched_scale_input(const float* x, float* x_in, int64_t n, float c_in, void* stream);
/* Heun prologue (edm_sampler.py:44-58): x = float64(noise) * t_0 ; x_in = float32(x * c_in(t_0)) */
LC_API int lc_sched_heun_init(const float* noise, double* x, float* x_in, int64_t n, double t0, double c_in, void* stream);
/* Heun stochastic churn (edm_sampler.py:67-76, deterministic=False): x += k * noise (fp64, k = sqrt(t_hat^2 - t_cur^2) *
 * S_noise) ; x_in = float32(x * c_in(t_hat)) */
LC_API int lc_sched_heun_churn(double* x, const double* noise, float* x_in, int64_t n, double k, double c_in, void* stream);
/* AR feedback of roll_out_serial (pipelines/utils.py:560-585) on one sampler output samples[B, C, T_out, hw]
 * (normalised latents): known_next[B, C, T_in, hw] = the last T_in frames (may be NULL); phys[B, C, T_out, hw] =
 * (samples / target_std) * std[c] + mean[c] (inverse_normalize_transform_3D, dataloader/utils.py:233-240; may be
 * NULL).  hw must be even. */
LC_API int lc_latent_feedback(const float* samples, float* known_next, float* phys, const float* mean, const float* std,
                              float target_std, int batch, int channels, int t_out, int t_in, int hw, void* stream);
/* phase 0: Euler predictor from x (saved to x_hat) ; phase 1: trapezoid corrector.  State in fp64. */
LC_API int lc_sched_heun_step(const float* f, double* x, double* x_hat, double* d_cur, float* x_in_next, int64_t n, int phase,
                       double t_cur, double t_next, double c_skip, double c_out, double c_in_next, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * DC-AE decoder — replaces AutoencoderDC.decode / Decoder.forward (models/DCAE.py:1018-1056, 717-732) and the
 * de-normalisation of decode_latent_ens (pipelines/utils.py:71-79).
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct lc_dcae lc_dcae;

typedef struct {
  int latent_channels; /* 84 */
  int out_channels;    /* decoder conv_out channels, 89 */
  int head_dim;        /* EfficientViT attention_head_dim, 32 */
  int n_stages;        /* len(decoder_block_out_channels), 4 */
  int precision;       /* LC_PRECISION_* */
  int stage_channels[8]; /* decoder_block_out_channels, index 0 = highest resolution: 252, 504, 504, 1008 */
  int stage_layers[8];   /* decoder_layers_per_block */
  int stage_is_evit[8];  /* 1 where decoder_block_types[i] == "EfficientViTBlock" */
  /* encoder (optional: leave in_channels = 0 for a decode-only handle); same number of stages as the decoder */
  int in_channels;           /* encoder conv_in channels = fields + static, 89 */
  int enc_stage_channels[8]; /* encoder_block_out_channels, index 0 = highest resolution */
  int enc_stage_layers[8];   /* encoder_layers_per_block (layers[0] must be > 0) */
  int enc_stage_is_evit[8];  /* 1 where encoder_block_types[i] == "EfficientViTBlock" */
} lc_dcae_cfg;

LC_API int lc_dcae_create(const lc_dcae_cfg* cfg, lc_dcae** out);
LC_API void lc_dcae_destroy(lc_dcae* h);
/* state-dict entries with prefix "decoder." and/or "encoder." (reference key names); the
 * decoder / encoder is built at finalize when its conv_in.weight was loaded */
LC_API int lc_dcae_load(lc_dcae* h, const char* key, const float* data, const int64_t* shape, int ndim, void* stream);
LC_API int lc_dcae_finalize(lc_dcae* h, void* stream);
/* allocates the workspace for up to max_frames latents of size h x w (the only allocating call after finalize) */
LC_API int lc_dcae_reserve(lc_dcae* h, int max_frames, int height, int width, void* stream);
/* z: [n, latent_channels, h, w] fp32 -> out: [n, keep_channels, 8h, 8w] fp32 (NCHW); keep_channels =
 * out_channels - static_channels (84) drops the static channels exactly like DCAE.py:1050-1052.  If mean/std
 * ([keep_channels] fp32) are given the output is out*std + mean (inverse_normalize_transform_3D,
 * dataloader/utils.py:233-240). */
LC_API int lc_dcae_decode(lc_dcae* h, const float* z, int n, int height, int width, float* out, int keep_channels,
                          const float* mean, const float* std, void* stream);

/* decode_latent_ens (pipelines/utils.py:52-80) without its permute / reshape copies: decodes frames
 * [frame0, frame0 + n) of the (batch x t_take) block — frame f = b * t_take + t is latents[b, :, t] of the 5-D tensor
 * latents [batch, latent_channels, t_total, h, w] read in place (t_take = extract_first <= t_total) — and writes it to
 * out[b, :, t] of out [batch, keep_channels, t_take, 8h, 8w].  lat_mean/lat_std ([latent_channels]) optionally apply
 * inverse_normalize_transform_3D to the latents first ((z / target_std) * std + mean, dataloader/utils.py:233-240, as
 * roll_out_serial does at pipelines/utils.py:571-577); mean/std as in lc_dcae_decode.  n <= max_frames of reserve. */
LC_API int lc_dcae_decode_ens(lc_dcae* h, const float* latents, int batch, int t_total, int t_take, int frame0, int n,
                              int height, int width, float* out, int keep_channels, const float* mean, const float* std,
                              const float* lat_mean, const float* lat_std, float target_std, void* stream);

/* x: [n, in_channels, 8h, 8w] fp32 NCHW (fields already concatenated with the static channels, DCAE.py:985-986)
 * -> out: [n, latent_channels, h, w] fp32 — replaces AutoencoderDC.encode / Encoder.forward (models/DCAE.py:964-1000,
 * 617-631), used once per forecast init time (pipelines/utils.py:471).  If mean/std ([latent_channels] fp32) are
 * given the output is (z - mean) / std * target_std (normalize_transform_3D, dataloader/utils.py:223-231). */
LC_API int lc_dcae_encode(lc_dcae* h, const float* x, int n, int height, int width, float* out, const float* mean,
                          const float* std, float target_std, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Ensemble metrics — replaces pointwise_crps_skill / pointwise_crps_spread / get_crps (evaluate/utils.py:52-118)
 * and the per-lead-time assembly of evaluate/evaluate_ens_gpu.py:339-415.
 * fields: [members, planes, H*W] fp32 (planes = (channel, lead) pairs), truth: [planes, H*W] fp32 (NaN allowed),
 * lat_weights: [H] fp64 (get_normalized_lat_weights_based_on_cos, evaluate/utils.py:40-48).
 * ---------------------------------------------------------------------------------------------------------- */
/* sums/counts: [4, planes] fp64 — latitude-weighted spatial SUMS and non-NaN pixel counts of
 * {ens-mean squared error, CRPS skill, CRPS spread, CRPS = skill - 0.5 spread}; mean = sum/count (nanmean) or
 * sum/(H*W) with NaN propagation is formed by the caller.  Zeroes the outputs itself. */
LC_API int lc_metrics_accumulate(const float* fields, const float* truth, const double* lat_weights, int members,
                                 long long planes, int height, int width, double* sums, double* counts, void* stream);
/* the same with an explicit distance (in elements) between consecutive members: member m's planes [planes, H*W] start at
 * fields + m * member_stride.  Reads a member-sharded receive buffer or a slice of a larger tensor in place. */
LC_API int lc_metrics_accumulate_strided(const float* fields, long long member_stride, const float* truth,
                                         const double* lat_weights, int members, long long planes, int height, int width,
                                         double* sums, double* counts, void* stream);
/* Peer-visible device buffers for the fused metrics exchange (one node, one process per GPU): the owner allocates
 * `bytes` of device memory and gets a 64-byte CUDA-IPC handle to hand to the other processes (any byte transport);
 * lc_ipc_open maps an owner's buffer for kernels of the CALLER's current GPU (cudaIpcOpenMemHandle with lazy peer
 * access: loads go over NVLink / NVSwitch).  Close every opened mapping before the owner frees. */
LC_API int lc_ipc_alloc(size_t bytes, void** dev_ptr, unsigned char* handle64);
LC_API int lc_ipc_open(const unsigned char* handle64, void** dev_ptr);
LC_API int lc_ipc_close(void* dev_ptr);
LC_API int lc_ipc_free(void* dev_ptr);
/* Lets kernels launched on the CURRENT device dereference memory of `peer_device` (cudaDeviceEnablePeerAccess;
 * already-enabled is not an error).  Needed once per peer before lc_metrics_accumulate_ptrs is given pointers into
 * another GPU's memory (e.g. CUDA-IPC mappings of the other ranks' decoded fields). */
LC_API int lc_enable_peer_access(int peer_device);
/* the same with one base pointer per member (HOST array of `members` <= 64 DEVICE pointers; member m's planes
 * [planes, H*W] are contiguous at member_ptrs[m]).  A pointer may address another GPU's memory mapped into this process
 * (CUDA IPC / peer access): with members sharded over GPUs (evaluate/evaluate_ens_gpu.py:462-468 gathers them instead)
 * every rank reduces its slice of the planes reading the other ranks' members in place over NVLink — the member->plane
 * exchange and the reduction are one kernel, no gathered copy. */
LC_API int lc_metrics_accumulate_ptrs(const float* const* member_ptrs, const float* truth, const double* lat_weights,
                                      int members, long long planes, int height, int width, double* sums, double* counts,
                                      void* stream);
/* anomaly-correlation terms of get_acc (evaluate/utils.py:122-149): sums/counts [3, planes] fp64 of the NaN-skipping
 * (optionally latitude-weighted) spatial sums of fa*ta, fa^2, ta^2 with fa = forecast - climate, ta = truth - climate */
LC_API int lc_metrics_acc(const float* forecast, const float* truth, const float* climate, const double* lat_weights,
                          long long planes, int height, int width, double* sums, double* counts, void* stream);
/* per-pixel outputs [planes, H*W] fp32 (any may be NULL): CRPS skill, CRPS spread, ensemble mean */
LC_API int lc_metrics_pointwise(const float* fields, const float* truth, int members, long long planes, int height,
                                int width, float* out_skill, float* out_spread, float* out_mean, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Low-level ops exported for parity tests (same kernels the handles use)
 * ---------------------------------------------------------------------------------------------------------- */
/* C[M,N] = A[M,K] W[N,K]^T + bias, act in {0 none, 1 gelu-tanh, 2 silu}.  precision BF16: A, W bf16 (raw uint16
 * storage), C fp32; precision F32: everything fp32. */
LC_API int lc_gemm(int precision, const void* a, const void* w, const float* bias, float* c, int m, int n, int k, int act,
            void* stream);
/* same product on the tensor-core path with a bf16 output matrix (the denoiser's activation storage type) */
LC_API int lc_gemm_bf16out(const void* a, const void* w, const float* bias, void* c, int m, int n, int k, int act,
                           void* stream);
/* qkv: [B, S, 3*heads*128] (q|k|v) fp32 (F32) or bf16 (BF16); out: [B, S, heads*128] same dtype. */
LC_API int lc_attention(int precision, const void* qkv, void* out, int batch, int seq, int heads, void* stream);

/* The index permutations of the path, exported so that the tests can assert them bit-exact:
 *  - lc_patchify: HunyuanVideoPatchEmbed's flatten(2).transpose(1,2) for patch (1,1,1) (models/embeddings.py:56-59):
 *    x [B, C, thw] fp32 -> tokens [B*thw, kp] (fp32 or bf16), token n = t*H*W + h*W + w, columns >= C zero;
 *  - lc_unpatchify_gemm: tokens [B*thw, k] x w[n_out, k]^T (+bias) stored channel-major as [B, n_out, thw] fp32 — the
 *    proj_out + unpatchify of LaDCast_3D_model.py:1047-1062 (patch size 1: out-feature f = channel c);
 *  - lc_pixel_(un)shuffle_shortcut: DCUpBlock2d / DCDownBlock2d tails (models/DCAE.py:519-536, 476-490) on NHWC fp32:
 *    shuffle: conv [n,H,W,4*cout], xin [n,H,W,cin] -> out [n,2H,2W,cout]; unshuffle: conv [n,H,W,cout/4],
 *    xin [n,H,W,cin] -> out [n,H/2,W/2,cout]. */
LC_API int lc_patchify(int precision, const float* x, void* tokens, int batch, int channels, int thw, int kp, void* stream);
LC_API int lc_unpatchify_gemm(int precision, const void* tokens, const void* w, const float* bias, float* out, int batch,
                              int thw, int n_out, int k, void* stream);
LC_API int lc_pixel_shuffle_shortcut(const float* conv, const float* xin, float* out, int n, int height, int width, int cin,
                                     int cout, void* stream);
LC_API int lc_pixel_unshuffle_shortcut(const float* conv, const float* xin, float* out, int n, int height, int width,
                                       int cin, int cout, void* stream);

/* LayerNorm (no affine, biased variance, eps) of x [rows, d] fp32 followed by the AdaLN modulation
 * y * (1 + scale[b, :]) + shift[b, :] with b = row / rows_per_sample (scale / shift rows `mod_stride` floats apart;
 * NULL = none), or by an affine w, b [d]; out is fp32 (F32) or bf16 (BF16).  The denoiser's norm1 / norm2 / norm_out
 * (diffusers AdaLayerNormZero / ZeroSingle / Continuous; LaDCast_3D_model.py:287-302, 524-552, 1044).  d % 128 == 0. */
LC_API int lc_layernorm_modulate(int precision, const float* x, void* out, int rows, int d, float eps, int rows_per_sample,
                                 const float* scale, const float* shift, int64_t mod_stride, const float* w,
                                 const float* b, void* stream);

/* One SphereConv2d 3x3 (models/sphere_conv.py:138-192), NCHW fp32 in/out, through the implicit-GEMM path.
 * w: [cout, cin, 3, 3], bias: [cout] or NULL.  Test helper: allocates and synchronises internally. */
LC_API int lc_sphere_conv3x3(int precision, const float* x, const float* w, const float* bias, float* out, int n, int cin,
                             int height, int width, int cout, int act, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LADCAST_B200_H_ */
