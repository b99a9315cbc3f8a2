import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools.microbench import bench_attn
bench_attn(20, 2250, 12)
bench_attn(20, 450, 12)
