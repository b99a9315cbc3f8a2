#!/bin/bash
# First-light run on a B200: each group in its own process so that a trapped kernel cannot poison later groups.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name" ; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n 25 gpurun_out/$name.log; }
run gemm_f32   python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "gemm and f32" -x
run gemm_bf16  python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "gemm and bf16"
run attn_f32   python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention and f32" -x
run attn_bf16  python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention and bf16"
run sched      python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "dpmpp2m"
run den_f32    python -m pytest tests/test_denoiser_gpu.py -q -m gpu -k "fp32 or samplers"
run den_bf16   python -m pytest tests/test_denoiser_gpu.py -q -m gpu -k "bf16 or independence"
