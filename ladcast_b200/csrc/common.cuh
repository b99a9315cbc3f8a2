// Shared declarations for the ladcast_b200 CUDA library (internal; the public C ABI is include/ladcast_b200.h).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <string>
#include <utility>

namespace lc {

// ---------------------------------------------------------------- errors (thread-local message, C-ABI style)
void set_error(const std::string& msg);
const char* last_error();

#define LC_CHECK_CUDA(expr)                                                                       \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      ::lc::set_error(std::string(#expr) + " -> " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + \
                      std::to_string(__LINE__) + ")");                                            \
      return -2;                                                                                  \
    }                                                                                             \
  } while (0)

#define LC_REQUIRE(cond, msg)                                                               \
  do {                                                                                      \
    if (!(cond)) {                                                                          \
      ::lc::set_error(std::string(msg) + " [" #cond "] (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
      return -1;                                                                            \
    }                                                                                       \
  } while (0)

// every kernel launch in the library is followed by exactly one LC_LAUNCH_CHECK(): error check + launch counter
void count_launch();
#define LC_LAUNCH_CHECK()              \
  do {                                 \
    ::lc::count_launch();              \
    LC_CHECK_CUDA(cudaGetLastError()); \
  } while (0)

#define LC_TRY(expr)          \
  do {                        \
    int _r = (expr);          \
    if (_r != 0) return _r;   \
  } while (0)

// ---------------------------------------------------------------- element types
typedef __nv_bfloat16 bf16;

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<bf16>(bf16 v) { return __bfloat162float(v); }

template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float gelu_tanh(float x) {
  // torch F.gelu(approximate="tanh"): 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3)))
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float u = k0 * (x + k1 * x * x * x);
  return 0.5f * x * (1.0f + tanhf(u));
}
__device__ __forceinline__ float silu(float x) { return x / (1.0f + __expf(-x)); }

enum Act { ACT_NONE = 0, ACT_GELU_TANH = 1, ACT_SILU = 2, ACT_RELU = 3 };
__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case ACT_GELU_TANH: return gelu_tanh(v);
    case ACT_SILU: return silu(v);
    case ACT_RELU: return fmaxf(v, 0.0f);
    default: return v;
  }
}

// ---------------------------------------------------------------- GEMM epilogue description
// C[M,N] = A[M,K] * W[N,K]^T (+bias) -> act -> one of the output modes.  Rows are "tokens"/"pixels"; a
// sample = `rows_per_sample` consecutive rows; per-sample vectors (gates) are indexed by row / rows_per_sample.
enum EpiMode {
  EPI_STORE = 0,       // out[orow, n] = act(acc + bias)                  (out dtype: bf16 or f32)
  EPI_GATED_RESID = 1, // out_f32[orow, n] += gate[b, n] * (acc + bias)   (gate may be null -> 1)
  EPI_UNPATCHIFY = 2,  // out_f32[b, n, row % rows_per_sample] = acc + bias   (token-major -> channel-major)
  EPI_RESID_STORE = 3, // out[orow, n] = act(acc + bias) + resid_f32[orow, n]  (conv shortcut adds)
  // DC-AE block tail fused into the producing conv / 1x1 GEMM (tensor-core path, N <= one tile so that a CTA holds whole
  // rows): y = acc * rsqrt(mean_n(acc^2) + norm_eps) * norm_w[n] + norm_b[n];  xres[row, n] += y (fp32 residual stream,
  // in place, row pitch ldr);  out[orow, n] = T(xres[row, n]) with the usual row remap (next conv's padded input).
  // Reference: ResBlock / EfficientViTBlock / GLUMBConv tails, models/DCAE.py:356-377, 254-262, 316-324.
  EPI_NORM_RESID = 4,
};

struct EpiParams {
  int mode = EPI_STORE;
  int act = ACT_NONE;
  int out_f32 = 0;            // EPI_STORE / EPI_RESID_STORE: 1 -> float output, 0 -> bf16 output
  const float* bias = nullptr;
  void* out = nullptr;
  long long ldo = 0;          // elements between consecutive output rows
  // output row remap: orow = (row / rows_per_sample) * out_rows_per_sample + out_row_offset + row % rows_per_sample
  int rows_per_sample = 1 << 30;
  int out_rows_per_sample = 1 << 30;
  int out_row_offset = 0;
  // optional second row segment (rows >= seg_rows; the single-stream blocks run pred and cond tokens in ONE launch):
  // r = row - seg_rows; orow = seg_out_base + (r / seg_rows_per_sample) * seg_out_rows_per_sample + seg_out_row_offset
  //                            + r % seg_rows_per_sample;  sample = r / seg_rows_per_sample.  0 = disabled.
  int seg_rows = 0;
  int seg_rows_per_sample = 1;
  int seg_out_rows_per_sample = 0;
  int seg_out_row_offset = 0;
  long long seg_out_base = 0;
  const float* gate = nullptr;  // [n_samples, gate_stride]
  long long gate_stride = 0;
  const float* resid = nullptr;  // EPI_RESID_STORE
  long long ldr = 0;             // row pitch of resid / xres
  float* xres = nullptr;         // EPI_NORM_RESID: fp32 residual stream updated in place (identity row map)
  const float* norm_w = nullptr; // EPI_NORM_RESID: RMSNorm weight / bias [N]
  const float* norm_b = nullptr;
  float norm_eps = 0.f;
  int stage_bf16 = 0;  // CTA-pair kernel, bf16 outputs: rows staged through shared memory, coalesced write-back (LADCAST_B200_EPI_STAGE=0: off)
  int prefetch = 0;  // read-modify-write epilogues: L2 prefetch of the next block's residual lines (LADCAST_B200_EPI_PREFETCH=1: on)
  int n_valid = 0;  // EPI_UNPATCHIFY: number of real output channels (<= N)
  // EPI_UNPATCHIFY into a 5-D [B, n_valid, up_T, rows_per_sample] tensor: sample s of this launch is frame
  // (up_frame0 + s) = b * up_T + t and lands in plane (b, n, t).  up_T = 1: plain [samples, n_valid, rows_per_sample].
  int up_T = 1;
  int up_frame0 = 0;
  // second-level remap (padded image buffers): orow += (row / rows_per_group) * group_extra_rows
  int rows_per_group = 1 << 30;
  int group_extra_rows = 0;
  // EPI_UNPATCHIFY only: out = v * ch_scale[n] + ch_shift[n]  (decoder: fields * std + mean fused)
  const float* ch_scale = nullptr;
  const float* ch_shift = nullptr;
  // fused per-head RMSNorm(q, k) + RoPE for qkv projections on the tensor-core path (head_dim 128): columns
  // [0, qk_cols) are q|k heads, normalised with qk_wq / qk_wk ([128] each) and rotated with the packed rope_cs
  // table (token = row % rows_per_sample; null = no rotation).  0 = disabled.
  int qk_cols = 0;
  float qk_eps = 0.f;
  const float* qk_wq = nullptr;
  const float* qk_wk = nullptr;
  const uint32_t* rope_cs = nullptr;  // [tokens, 64] half2 (cos, sin) per rotation pair; null = no rotation
};

__device__ __forceinline__ long long epi_out_row(const EpiParams& ep, int row, int& sample) {
  if (ep.seg_rows > 0 && row >= ep.seg_rows) {
    const int r2 = row - ep.seg_rows;
    sample = r2 / ep.seg_rows_per_sample;
    return ep.seg_out_base + static_cast<long long>(sample) * ep.seg_out_rows_per_sample + ep.seg_out_row_offset +
           (r2 - sample * ep.seg_rows_per_sample);
  }
  sample = row / ep.rows_per_sample;
  int r = row - sample * ep.rows_per_sample;
  return static_cast<long long>(sample) * ep.out_rows_per_sample + ep.out_row_offset + r +
         static_cast<long long>(row / ep.rows_per_group) * ep.group_extra_rows;
}

// element offset of (sample, channel 0, row-in-sample 0) and the channel stride of an EPI_UNPATCHIFY output
__device__ __forceinline__ long long epi_unpatch_base(const EpiParams& ep, int sample, long long& ch_stride) {
  const int fg = sample + ep.up_frame0;
  const int b = fg / ep.up_T, t = fg - b * ep.up_T;
  ch_stride = static_cast<long long>(ep.up_T) * ep.rows_per_sample;
  return (static_cast<long long>(b) * ep.n_valid * ep.up_T + t) * ep.rows_per_sample;
}

// GEMM problem: A is [M, K] row-major split in up to two K-segments (A0: k < K0, A1: K0 <= k < K);
// W is [N, K] row-major (nn.Linear layout).  bf16 path: A/W bf16, fp32 path: A/W float.
struct GemmArgs {
  const void* A0 = nullptr;
  long long lda0 = 0;
  int K0 = 0;
  const void* A1 = nullptr;
  long long lda1 = 0;
  const void* W = nullptr;
  long long ldw = 0;
  int M = 0, N = 0, K = 0;
  EpiParams epi;
};

int gemm_f32(const GemmArgs& g, cudaStream_t stream);   // SIMT fp32 validation path
int gemm_bf16(const GemmArgs& g, cudaStream_t stream);
int gemm_set_trace(long long* buf);  // tuning aid: timeline of CTA pair 0 of gemm_tc2_kernel (nullptr = off)  // tcgen05 / TMEM / TMA path
int gemm_bf16_selftest_smem_bytes();

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

int num_sms();  // of the CURRENT device (cached per device)

// Function attributes (cudaFuncSetAttribute) and device properties are per device, not per process: one-time guards
// are indexed by the current device so that a process driving several GPUs opts every one of them in.
constexpr int LC_MAX_DEVICES = 64;
inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < LC_MAX_DEVICES) ? dev : 0;
}
template <typename V>
struct PerDevice {
  V v[LC_MAX_DEVICES] = {};
  V& here() { return v[current_device()]; }
};

// ---------------------------------------------------------------- programmatic dependent launch (PDL)
// Every hot-path kernel calls pdl_grid_sync() after its block-local setup (barrier init, TMEM allocation, tensor-map
// prefetch, index math) and BEFORE its first global-memory access: griddepcontrol.wait blocks until the preceding grid
// of the stream has completed and flushed, then griddepcontrol.launch_dependents lets the NEXT grid's CTAs be
// scheduled as soon as every CTA of this grid is resident — so a kernel's launch latency and prologue overlap the
// tail of its predecessor while all memory ordering stays that of the stream.  (Triggering only after the wait keeps
// at most one dependent grid waiting on the SMs.)  Both instructions are no-ops for a grid launched without the
// attribute; launch_kernel() sets it when LADCAST_B200_PDL=1.
__device__ __forceinline__ void pdl_grid_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// ---------------------------------------------------------------- lightweight per-class kernel timing (bench only)
// When enabled, every launch of the library is bracketed with CUDA events on its stream; durations, algorithmic FLOPs
// and algorithmic bytes are accumulated per class at collect time (bench.py: roofline + roofline.secondary).
// Disabled by default (zero overhead: one branch per launch).
enum ProfClass {
  PROF_GEMM = 0,       // tcgen05 GEMM (denoiser linears, decoder 1x1)
  PROF_ATTN = 1,       // tcgen05 flash attention
  PROF_CONV = 2,       // implicit-GEMM 3x3 sphere convolution
  PROF_LN = 3,         // LayerNorm + modulation
  PROF_ROPE = 4,       // per-head RMSNorm(q, k) + RoPE
  PROF_SCHED = 5,      // scheduler steps / latent feedback
  PROF_DEC_NORM = 6,   // decoder channel RMSNorm (+ residual)
  PROF_DEC_MS = 7,     // decoder fused multiscale projection
  PROF_DEC_LINATTN = 8,// decoder ReLU linear attention
  PROF_DEC_DWGLU = 9,  // decoder depthwise 3x3 + GLU
  PROF_DEC_SHUFFLE = 10,  // pixel (un)shuffle + shortcut
  PROF_DEC_PAD = 11,   // sphere padding / halo fill / shortcuts
  PROF_METRICS = 12,   // ensemble metrics reduction
  PROF_MISC = 13,      // patchify, embeddings, pooling, casts
  PROF_NUM = 14
};
const char* prof_name(int cls);
bool prof_on();
void prof_begin(int cls, cudaStream_t s);
void prof_end(int cls, double flops, cudaStream_t s, double bytes = 0.0);
// RAII bracket around one launch: ProfScope ps(PROF_LN, 0, bytes, stream); kernel<<<...>>>(...);
struct ProfScope {
  int cls;
  double flops, bytes;
  cudaStream_t s;
  ProfScope(int c, double fl, double by, cudaStream_t st) : cls(c), flops(fl), bytes(by), s(st) { prof_begin(cls, s); }
  ~ProfScope() { prof_end(cls, flops, s, bytes); }
};

}  // namespace lc
