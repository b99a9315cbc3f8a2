#!/bin/bash
# A/B: block-ahead L2 prefetch in the read-modify-write GEMM epilogues (denoiser call time), then parity of the build.
set -u
mkdir -p gpurun_out
for i in 1 2; do for pf in 1 0; do
  echo "prefetch=$pf"; LADCAST_B200_EPI_PREFETCH=$pf timeout 300 python tools/bench_den.py 2>&1 | grep "denoiser"
done; done
echo "=== pytest (kernels + denoiser)"; timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_denoiser_gpu.py -m gpu -q -x 2>&1 | tail -2
