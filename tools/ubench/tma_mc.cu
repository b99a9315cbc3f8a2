// Micro-benchmark: L2 -> shared-memory delivery rate of TMA tile loads when every CTA of a cluster needs the SAME
// 16 KB tile: unicast (each CTA loads it) vs multicast (each CTA loads 1/C of it and broadcasts).  Answers whether
// operand multicast over clusters of C CTAs lifts the ~6.3 kB/clk L2 cap the pair GEMM sits on.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -I ladcast_b200/csrc tools/ubench/tma_mc.cu ladcast_b200/csrc/tmap.cu
#include <cstdio>
#include <cstdint>
#include <string>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "ptx.cuh"
#include "tmap.h"
namespace lc { void set_error(const std::string& m) { fprintf(stderr, "error: %s\n", m.c_str()); } }
using namespace lc;

constexpr int STAGES = 8;
constexpr int TILE_ROWS = 128;                       // [128 rows][64 bf16] = 16 KB
constexpr int TILE_BYTES = TILE_ROWS * 64 * 2;

__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(ptx::smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}

template <int MC>
__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap tm, int csz, int iters, int rows_total,
                                            long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  __shared__ uint64_t full[STAGES];
  const uint32_t rank = ptx::cluster_ctarank();
  const int cluster_id = blockIdx.x / csz;
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) ptx::mbar_init(&full[i], 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  ptx::cluster_sync();
  long long t0 = clock64();
  uint32_t ph = 0;
  const int slice = TILE_ROWS / csz;
  for (int it = 0; it < iters; ++it) {
    if (threadIdx.x == 0) {
      for (int s = 0; s < STAGES; ++s) {
        // a different tile every time, the same one for every CTA of the cluster
        const int row0 = ((cluster_id * 131 + it * STAGES + s) * TILE_ROWS) % (rows_total - TILE_ROWS);
        ptx::mbar_expect_tx(&full[s], TILE_BYTES);
        if (MC) tma_load_2d_mc(smem + s * TILE_BYTES + rank * slice * 128, &tm, &full[s], 0, row0 + rank * slice,
                               static_cast<uint16_t>((1u << csz) - 1));
        else ptx::tma_load_2d(smem + s * TILE_BYTES, &tm, &full[s], 0, row0);
      }
      for (int s = 0; s < STAGES; ++s) ptx::mbar_wait_spin(&full[s], ph);
    }
    ph ^= 1;
    __syncthreads();
    ptx::cluster_sync();  // nobody refills a slot that a peer's multicast may still be writing
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

template <int MC>
void run(const CUtensorMap& tm, int csz, int rows_total, long long* d_out) {
  const int smem = STAGES * TILE_BYTES + 2048;
  cudaFuncSetAttribute(k<MC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k<MC>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  const int iters = 400;
  int grid = 148 / csz * csz;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = csz; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int max_clusters = 0;
  cudaOccupancyMaxActiveClusters(&max_clusters, k<MC>, &cfg);
  cudaError_t e = cudaLaunchKernelEx(&cfg, k<MC>, tm, csz, iters, rows_total, d_out);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("csz %d MC %d: %s\n", csz, MC, cudaGetErrorString(e)); return; }
  long long h[148];
  cudaMemcpy(h, d_out, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  const double bytes = double(iters) * STAGES * TILE_BYTES;  // delivered per CTA
  printf("cluster %d %-9s: %6.1f B/clk/SM delivered, %7.0f B/clk chip-wide (%d CTAs, max co-resident clusters %d)\n", csz,
         MC ? "multicast" : "unicast", bytes / mx, bytes / mx * grid, grid, max_clusters);
}

int main() {
  const int rows_total = 1 << 20;  // 128 MB bf16 [rows][64]: L2-resident after the first touches
  __nv_bfloat16* buf;
  cudaMalloc(&buf, size_t(rows_total) * 64 * 2);
  cudaMemset(buf, 0, size_t(rows_total) * 64 * 2);
  long long* d_out;
  cudaMalloc(&d_out, 148 * 8);
  for (int csz : {1, 2, 4, 8}) {
    CUtensorMap tm_uc, tm_mc;
    make_tmap_2d_bf16(&tm_uc, buf, 64, rows_total, 128, 64, TILE_ROWS);
    make_tmap_2d_bf16(&tm_mc, buf, 64, rows_total, 128, 64, TILE_ROWS / csz);
    run<0>(tm_uc, csz, rows_total, d_out);
    if (csz > 1) run<1>(tm_mc, csz, rows_total, d_out);
  }
  return 0;
}
