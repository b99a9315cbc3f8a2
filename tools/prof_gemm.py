"""Two GEMM launches for ncu source-level profiling: f32-out and bf16-out (GELU) epilogues."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ladcast_b200 import _lib
lib = _lib.load()
m, n, k = 36000, 4608, 1536
a = torch.randn(m, k, device="cuda").bfloat16(); w = torch.randn(n, k, device="cuda").bfloat16(); b = torch.randn(n, device="cuda")
c32 = torch.empty(m, n, device="cuda"); c16 = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    _lib.check(lib.lc_gemm(0, _lib.ptr(a), _lib.ptr(w), _lib.ptr(b), _lib.ptr(c32), m, n, k, 0, _lib.stream()))
    _lib.check(lib.lc_gemm_bf16out(_lib.ptr(a), _lib.ptr(w), _lib.ptr(b), _lib.ptr(c16), m, n, k, 1, _lib.stream()))
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e[0].record(); _lib.check(lib.lc_gemm(0, _lib.ptr(a), _lib.ptr(w), _lib.ptr(b), _lib.ptr(c32), m, n, k, 0, _lib.stream()))
e[1].record(); _lib.check(lib.lc_gemm_bf16out(_lib.ptr(a), _lib.ptr(w), _lib.ptr(b), _lib.ptr(c16), m, n, k, 1, _lib.stream()))
e[2].record(); torch.cuda.synchronize()
print("f32 out ms", e[0].elapsed_time(e[1]), "bf16 gelu out ms", e[1].elapsed_time(e[2]))
