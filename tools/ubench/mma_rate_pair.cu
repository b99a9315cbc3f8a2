// Micro-benchmark: issue rate of tcgen05.mma.cta_group::2 (M=256 across a CTA pair) for attention-like shapes.
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace lc;

// MODE 0: A smem K-major, B K-major (each CTA holds N/2 rows of B);  MODE 1: B MN-major;  A_TMEM: A from tensor memory
template <int N, int MODE, int A_TMEM>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) rate_kernel(long long* out, int iters, int nmma) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t rank = ptx::cluster_ctarank();
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
  if (threadIdx.x < 32) ptx::tmem_alloc_pair<512>(&slot);
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tm = slot;
  if (rank == 0 && threadIdx.x == 0) {
    const uint32_t a_addr = ptx::smem_u32(smem), b_addr = a_addr + 64 * 1024;
    constexpr uint32_t idesc = ptx::make_idesc_bf16(256, N, 0, MODE == 1 ? 1 : 0);
    uint32_t ph = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      for (int kk = 0; kk < nmma; ++kk) {
        const uint32_t k8 = kk & 7;
        uint64_t da = ptx::make_smem_desc(a_addr + (k8 >> 2) * 16384 + (k8 & 3) * 32, 16, 1024);
        uint64_t db;
        if (MODE == 0) db = ptx::make_smem_desc(b_addr + (k8 >> 2) * (N / 2 * 128) + (k8 & 3) * 32, 16, 1024);
        else db = ptx::make_smem_desc(b_addr + (k8 & 3) * 2048, 8192, 1024);
        if (A_TMEM) {
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tm + (it & 1) * 128),
              "r"(tm + 384 + k8 * 8), "l"(db), "r"(idesc), "r"(kk != 0 ? 1u : 0u)
              : "memory");
        } else {
          ptx::umma_f16_pair(tm + (it & 1) * 128, da, db, idesc, kk != 0 ? 1u : 0u);
        }
      }
      ptx::umma_commit_pair(&bar);
      ptx::mbar_wait_spin(&bar, ph);
      ph ^= 1;
    }
    long long t1 = clock64();
    out[blockIdx.x / 2] = t1 - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  if (threadIdx.x < 32) ptx::tmem_dealloc_pair<512>(tm);
}

template <int N, int MODE, int A_TMEM>
void run(const char* name, int nmma, int grid) {
  long long* d;
  cudaMalloc(&d, sizeof(long long) * grid);
  const int smem = 161 * 1024 + 1024;
  cudaFuncSetAttribute(rate_kernel<N, MODE, A_TMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 200;
  rate_kernel<N, MODE, A_TMEM><<<grid, 128, smem>>>(d, iters, nmma);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  long long h;
  cudaMemcpy(&h, d, sizeof(long long), cudaMemcpyDeviceToHost);
  double cyc = double(h) / iters;
  double mac = 256.0 * N * 16 * nmma;
  printf("%-48s grid %3d nmma/commit %3d: %8.1f cyc/batch %7.1f cyc/mma %7.0f MAC/clk per SM\n", name, grid, nmma, cyc,
         cyc / nmma, mac / cyc / 2);
  cudaFree(d);
}

int main() {
  for (int nmma : {8, 64}) {
    run<128, 0, 0>("pair N=128 A smem, B K-major (QK, 128 keys)", nmma, 148);
    run<256, 0, 0>("pair N=256 A smem, B K-major (QK, 256 keys)", nmma, 148);
    run<128, 1, 0>("pair N=128 A smem, B MN-major (PV)", nmma, 148);
    run<128, 1, 1>("pair N=128 A tmem, B MN-major (PV, P in TMEM)", nmma, 148);
  }
  return 0;
}
