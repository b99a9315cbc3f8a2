// Remaining C-ABI entry points: version/error, scheduler steps, low-level op exports used by the parity tests.
#include <cstring>
#include <string>

#include "../../include/ladcast_b200.h"
#include "kernels.h"

#include <atomic>
#include <cstdlib>
#include <vector>

namespace lc {
static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
const char* last_error() { return g_err.c_str(); }

bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("LADCAST_B200_PDL"); return e != nullptr && e[0] == '1'; }();
  return on;
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

struct ProfRec { cudaEvent_t a, b; int cls; double flops, bytes; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static cudaEvent_t g_pending[PROF_NUM];
static const char* const g_prof_names[PROF_NUM] = {
    "gemm_tc", "attention_tc", "sphere_conv_tc", "layernorm", "qk_norm_rope", "scheduler", "dec_rmsnorm",
    "dec_multiscale", "dec_linear_attn", "dec_dwconv_glu", "dec_pixel_shuffle", "dec_pad", "metrics", "misc"};
const char* prof_name(int cls) { return (cls >= 0 && cls < PROF_NUM) ? g_prof_names[cls] : "?"; }
bool prof_on() { return g_prof_on; }
void prof_begin(int cls, cudaStream_t s) {
  if (!g_prof_on) return;
  cudaEvent_t e;
  cudaEventCreate(&e);
  cudaEventRecord(e, s);
  g_pending[cls] = e;
}
void prof_end(int cls, double flops, cudaStream_t s, double bytes) {
  if (!g_prof_on) return;
  ProfRec r;
  r.a = g_pending[cls];
  cudaEventCreate(&r.b);
  cudaEventRecord(r.b, s);
  r.cls = cls;
  r.flops = flops;
  r.bytes = bytes;
  g_prof.push_back(r);
}
}  // namespace lc

using namespace lc;

extern "C" {

int lc_version(void) { return 100; }

long long lc_launch_count(void) { return lc::g_launches.load(); }

int lc_prof_enable(int on) {
  for (auto& r : lc::g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  lc::g_prof.clear();
  lc::g_prof_on = on != 0;
  return 0;
}

int lc_prof_collect(double* ms, double* flops, long long* launches) {
  LC_REQUIRE(ms && flops && launches, "null argument");
  for (int i = 0; i < 3; ++i) { ms[i] = 0; flops[i] = 0; launches[i] = 0; }
  for (auto& r : lc::g_prof) {
    if (r.cls >= 3) continue;
    LC_CHECK_CUDA(cudaEventSynchronize(r.b));
    float t = 0.f;
    LC_CHECK_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
    ms[r.cls] += t;
    flops[r.cls] += r.flops;
    launches[r.cls] += 1;
  }
  return 0;
}
int lc_prof_num_classes(void) { return PROF_NUM; }
const char* lc_prof_class_name(int cls) { return lc::prof_name(cls); }
int lc_prof_collect_all(double* ms, double* flops, double* bytes, long long* launches) {
  LC_REQUIRE(ms && flops && bytes && launches, "null argument");
  for (int i = 0; i < PROF_NUM; ++i) { ms[i] = 0; flops[i] = 0; bytes[i] = 0; launches[i] = 0; }
  for (auto& r : lc::g_prof) {
    LC_CHECK_CUDA(cudaEventSynchronize(r.b));
    float t = 0.f;
    LC_CHECK_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
    ms[r.cls] += t;
    flops[r.cls] += r.flops;
    bytes[r.cls] += r.bytes;
    launches[r.cls] += 1;
  }
  return 0;
}
const char* lc_last_error(void) { return lc::last_error(); }

int lc_sched_dpmpp2m_step(const float* f, float* x, float* x0_prev, float* x_in_next, int64_t n, float c_skip,
                          float c_out, float a_x, float a_x0, float a_d, float c_in_next, void* stream) {
  LC_REQUIRE(f && x && x0_prev, "null argument");
  SchedCoef c;
  c.c_skip = c_skip; c.c_out = c_out; c.a_x = a_x; c.a_x0 = a_x0; c.a_d = a_d; c.c_in_next = c_in_next;
  return sched_dpmpp2m_step(f, x, x0_prev, x_in_next, n, c, static_cast<cudaStream_t>(stream));
}

int lc_sched_heun_step(const float* f, double* x, double* x_hat, double* d_cur, float* x_in_next, int64_t n, int phase,
                       double t_cur, double t_next, double c_skip, double c_out, double c_in_next, void* stream) {
  LC_REQUIRE(f && x && x_hat && d_cur, "null argument");
  return sched_heun_step(f, x, x_hat, d_cur, x_in_next, n, phase, t_cur, t_next, c_skip, c_out, c_in_next,
                         static_cast<cudaStream_t>(stream));
}

int lc_sched_scale_input(const float* x, float* x_in, int64_t n, float c_in, void* stream) {
  LC_REQUIRE(x && x_in, "null argument");
  return sched_scale_input(x, x_in, n, c_in, static_cast<cudaStream_t>(stream));
}

int lc_sched_heun_init(const float* noise, double* x, float* x_in, int64_t n, double t0, double c_in, void* stream) {
  LC_REQUIRE(noise && x && x_in, "null argument");
  return sched_heun_init(noise, x, x_in, n, t0, c_in, static_cast<cudaStream_t>(stream));
}

int lc_sched_heun_churn(double* x, const double* noise, float* x_in, int64_t n, double k, double c_in, void* stream) {
  LC_REQUIRE(x && noise && x_in, "null argument");
  return sched_heun_churn(x, noise, x_in, n, k, c_in, static_cast<cudaStream_t>(stream));
}

// ---- peer-visible device buffers (CUDA IPC): the owner allocates and exports, every other process of the node opens
// the handle ON ITS OWN GPU with lazy peer access, which maps the owner's memory for this GPU's kernels over NVLink.
int lc_ipc_alloc(size_t bytes, void** dev_ptr, unsigned char* handle64) {
  LC_REQUIRE(dev_ptr != nullptr && handle64 != nullptr && bytes > 0, "bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  LC_CHECK_CUDA(cudaMalloc(&p, bytes));
  cudaIpcMemHandle_t h;
  const cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    LC_CHECK_CUDA(e);
  }
  memcpy(handle64, &h, 64);
  *dev_ptr = p;
  return 0;
}

int lc_ipc_open(const unsigned char* handle64, void** dev_ptr) {
  LC_REQUIRE(dev_ptr != nullptr && handle64 != nullptr, "bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  LC_CHECK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *dev_ptr = p;
  return 0;
}

int lc_ipc_close(void* dev_ptr) {
  if (dev_ptr != nullptr) LC_CHECK_CUDA(cudaIpcCloseMemHandle(dev_ptr));
  return 0;
}

int lc_ipc_free(void* dev_ptr) {
  if (dev_ptr != nullptr) LC_CHECK_CUDA(cudaFree(dev_ptr));
  return 0;
}

int lc_enable_peer_access(int peer_device) {
  int dev = 0;
  LC_CHECK_CUDA(cudaGetDevice(&dev));
  if (peer_device == dev) return 0;
  int can = 0;
  LC_CHECK_CUDA(cudaDeviceCanAccessPeer(&can, dev, peer_device));
  LC_REQUIRE(can != 0, "the current GPU cannot access the requested peer GPU (no NVLink / PCIe peer path)");
  const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    (void)cudaGetLastError();  // not an error: somebody (torch, NCCL) enabled it before
    return 0;
  }
  LC_CHECK_CUDA(e);
  return 0;
}

int lc_latent_feedback(const float* samples, float* known_next, float* phys, const float* mean, const float* stdv,
                       float target_std, int batch, int channels, int t_out, int t_in, int hw, void* stream) {
  LC_REQUIRE(samples && (known_next || phys), "null argument");
  LC_REQUIRE(batch > 0 && channels > 0 && t_out > 0 && hw > 0, "bad shape");
  return latent_feedback(samples, known_next, phys, mean, stdv, target_std, batch, channels, t_out, t_in, hw,
                         static_cast<cudaStream_t>(stream));
}

int lc_gemm(int precision, const void* a, const void* w, const float* bias, float* c, int m, int n, int k, int act,
            void* stream) {
  LC_REQUIRE(a && w && c, "null argument");
  GemmArgs g;
  g.A0 = a; g.lda0 = k; g.K0 = k; g.W = w; g.ldw = k; g.M = m; g.N = n; g.K = k;
  g.epi.mode = EPI_STORE; g.epi.act = act; g.epi.bias = bias; g.epi.out = c; g.epi.ldo = n; g.epi.out_f32 = 1;
  return precision == LC_PRECISION_F32 ? gemm_f32(g, static_cast<cudaStream_t>(stream))
                                       : gemm_bf16(g, static_cast<cudaStream_t>(stream));
}

int lc_gemm_bf16out(const void* a, const void* w, const float* bias, void* c, int m, int n, int k, int act, void* stream) {
  LC_REQUIRE(a && w && c, "null argument");
  GemmArgs g;
  g.A0 = a; g.lda0 = k; g.K0 = k; g.W = w; g.ldw = k; g.M = m; g.N = n; g.K = k;
  g.epi.mode = EPI_STORE; g.epi.act = act; g.epi.bias = bias; g.epi.out = c; g.epi.ldo = n; g.epi.out_f32 = 0;
  return gemm_bf16(g, static_cast<cudaStream_t>(stream));
}

// Test exports of the two index permutations of the denoiser (north star: "index/patch permutations bit-exact").
int lc_patchify(int precision, const float* x, void* tokens, int batch, int channels, int thw, int kp, void* stream) {
  LC_REQUIRE(x && tokens && kp >= channels, "bad patchify arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return precision == LC_PRECISION_F32 ? patchify<float>(x, reinterpret_cast<float*>(tokens), batch, channels, thw, kp, st)
                                       : patchify<bf16>(x, reinterpret_cast<bf16*>(tokens), batch, channels, thw, kp, st);
}

int lc_unpatchify_gemm(int precision, const void* tokens, const void* w, const float* bias, float* out, int batch, int thw,
                       int n_out, int k, void* stream) {
  LC_REQUIRE(tokens && w && out, "null argument");
  GemmArgs g;
  g.A0 = tokens; g.lda0 = k; g.K0 = k; g.W = w; g.ldw = k; g.M = batch * thw; g.N = n_out; g.K = k;
  g.epi.mode = EPI_UNPATCHIFY; g.epi.bias = bias; g.epi.out = out; g.epi.rows_per_sample = thw; g.epi.n_valid = n_out;
  return precision == LC_PRECISION_F32 ? gemm_f32(g, static_cast<cudaStream_t>(stream))
                                       : gemm_bf16(g, static_cast<cudaStream_t>(stream));
}

// Test / micro-benchmark export of the LayerNorm + modulation kernel (AdaLayerNormZero / Continuous: LN without affine,
// eps, then y * (1 + scale[b]) + shift[b]; or an affine LN when w / b are given).
int lc_layernorm_modulate(int precision, const float* x, void* out, int rows, int d, float eps, int rows_per_sample,
                          const float* scale, const float* shift, int64_t mod_stride, const float* w, const float* b,
                          void* stream) {
  LC_REQUIRE(x && out, "null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return precision == LC_PRECISION_F32
             ? layernorm_modulate<float>(x, reinterpret_cast<float*>(out), rows, d, eps, rows_per_sample, scale, shift, mod_stride, w, b, st)
             : layernorm_modulate<bf16>(x, reinterpret_cast<bf16*>(out), rows, d, eps, rows_per_sample, scale, shift, mod_stride, w, b, st);
}

int lc_debug_gemm_trace(void* buf) { return lc::gemm_set_trace(reinterpret_cast<long long*>(buf)); }
int lc_debug_attention_trace(void* buf) { return lc::attention_set_trace(reinterpret_cast<long long*>(buf)); }
int lc_attention(int precision, const void* qkv, void* out, int batch, int seq, int heads, void* stream) {
  LC_REQUIRE(qkv && out, "null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (precision == LC_PRECISION_F32)
    return attention_f32(reinterpret_cast<const float*>(qkv), batch, seq, heads, 128, reinterpret_cast<float*>(out), seq,
                         nullptr, st);
  return attention_bf16(reinterpret_cast<const bf16*>(qkv), batch, seq, heads, 128, reinterpret_cast<bf16*>(out), seq,
                        nullptr, st);
}

}  // extern "C"
