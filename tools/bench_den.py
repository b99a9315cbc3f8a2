import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools.microbench import bench_gemm, bench_denoiser
bench_gemm(36000, 1536, 1536)
bench_gemm(36000, 1536, 6144)
bench_denoiser("375M", 20)
