import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools.microbench import bench_gemm
for shp in [(36000, 4608, 1536), (36000, 1536, 6144)]:
    bench_gemm(*shp)
