"""Timeline of CTA pair 0 of the pair GEMM: MMA issuer vs epilogue warp 0, per tile (cycles)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ladcast_b200 import _lib
lib = _lib.load()
lib.lc_debug_gemm_trace.argtypes = [ctypes.c_void_p]
m, n, k = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (36000, 6144, 1536)
act = int(sys.argv[4]) if len(sys.argv) > 4 else 1  # 1 = GELU-tanh
a = torch.randn(m, k, device="cuda").bfloat16(); w = torch.randn(n, k, device="cuda").bfloat16()
bias = torch.randn(n, device="cuda"); c = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
run = lambda: _lib.check(lib.lc_gemm_bf16out(_lib.ptr(a), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(c), m, n, k, act, _lib.stream()))
for _ in range(3): run()
tr = torch.zeros(128, 4, dtype=torch.int64, device="cuda")
lib.lc_debug_gemm_trace(_lib.ptr(tr)); run(); torch.cuda.synchronize(); lib.lc_debug_gemm_trace(None)
t = tr.cpu()
t0 = int(t[0, 0])
print(f"GEMM M={m} N={n} K={k} act={act}")
print("tile | MMA: start acc_free first_full issued(+commit) | EPI: wait accumulator_ready drained")
prev_issue = None
for i in range(40):
    if int(t[i, 3]) == 0: break
    mm = [int(x) - t0 for x in t[i]]; ee = [int(x) - t0 for x in t[64 + i][:3]]
    print(f"{i:3d} | {mm[0]:8d} {mm[1]:8d} {mm[2]:8d} {mm[3]:8d} | {ee[0]:8d} {ee[1]:8d} {ee[2]:8d}   mainloop issue {mm[3]-mm[1]:6d}  epilogue {ee[2]-ee[1]:6d}  acc wait {mm[1]-mm[0]:6d}")
