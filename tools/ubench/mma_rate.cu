// Micro-benchmark: issue rate of tcgen05.mma (kind::f16, cta_group::1, M=128) for the operand shapes the attention
// kernel uses.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -I ladcast_b200/csrc -o tools/ubench/mma_rate ...
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace lc;

// mode 0: A K-major 128x16 steps from a [128][64] swizzled tile; B K-major [N][64]; 4 k-steps per "tile"
// mode 1: B MN-major (V-like): [64 keys][64 dims] boxes, N = 128 -> two boxes (LBO = 8 KB)
template <int N, int MODE, int A_TMEM>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int iters, int nmma) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
  if (threadIdx.x < 32) ptx::tmem_alloc<512>(&slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tm = slot;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  ptx::fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t a_addr = ptx::smem_u32(smem), b_addr = a_addr + 64 * 1024;
    constexpr uint32_t idesc = ptx::make_idesc_bf16(128, N, 0, MODE == 1 ? 1 : 0);
    uint32_t ph = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      for (int kk = 0; kk < nmma; ++kk) {
        const uint32_t k8 = kk & 7;
        uint64_t da = ptx::make_smem_desc(a_addr + (k8 >> 2) * 16384 + (k8 & 3) * 32, 16, 1024);
        uint64_t db;
        if (MODE == 0) db = ptx::make_smem_desc(b_addr + (k8 >> 2) * (N * 128) + (k8 & 3) * 32, 16, 1024);
        else db = ptx::make_smem_desc(b_addr + (k8 & 3) * 2048, 8192, 1024);
        if (A_TMEM) {
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tm + (it & 1) * 256),
              "r"(tm + 384 + k8 * 8), "l"(db), "r"(idesc), "r"(kk != 0 ? 1u : 0u)
              : "memory");
        } else {
          ptx::umma_f16(tm + (it & 1) * 256, da, db, idesc, kk != 0 ? 1u : 0u);
        }
      }
      ptx::umma_commit(&bar);
      ptx::mbar_wait_spin(&bar, ph);
      ph ^= 1;
    }
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc<512>(tm);
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
  long long h[148];
  cudaMemcpy(h, d, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  double cyc = double(h[0]) / iters;
  double mac = 128.0 * N * 16 * nmma;
  printf("%-44s grid %3d nmma/commit %3d: %8.1f cyc/batch %7.1f cyc/mma %7.0f MAC/clk\n", name, grid, nmma, cyc, cyc / nmma,
         mac / cyc);
  cudaFree(d);
}

int main() {
  for (int grid : {1, 148}) {
    for (int nmma : {8, 64}) {
      run<64, 0, 0>("N=64  A smem K-major, B K-major (QK now)", nmma, grid);
      run<128, 0, 0>("N=128 A smem K-major, B K-major (QK BKV=128)", nmma, grid);
      run<256, 0, 0>("N=256 A smem K-major, B K-major", nmma, grid);
      run<128, 1, 0>("N=128 A smem, B MN-major (PV now)", nmma, grid);
      run<128, 1, 1>("N=128 A tmem, B MN-major (PV, P in TMEM)", nmma, grid);
      run<128, 0, 1>("N=128 A tmem, B K-major", nmma, grid);
    }
  }
  return 0;
}
