"""Timeline of attention CTA (0,0,0): where the tensor-pipe issuer and the two softmax tiles spend their cycles."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ladcast_b200 import _lib
lib = _lib.load()
b, s, h = 20, 2250, 12
d = h * 128
qkv = torch.randn(b, s, 3 * d, device="cuda").bfloat16()
out = torch.empty(b, s, d, device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    _lib.check(lib.lc_attention(0, _lib.ptr(qkv), _lib.ptr(out), b, s, h, _lib.stream()))
tr = torch.zeros(4, 64, 4, dtype=torch.int64, device="cuda")
lib.lc_debug_attention_trace.argtypes = [ctypes.c_void_p]
lib.lc_debug_attention_trace(_lib.ptr(tr))
_lib.check(lib.lc_attention(0, _lib.ptr(qkv), _lib.ptr(out), b, s, h, _lib.stream()))
torch.cuda.synchronize()
lib.lc_debug_attention_trace(None)
t = tr.cpu()
n = (s + 127) // 128
t0 = int(t[1, 0, 0])
print("step | MMA: pfullA issuedA pfullB issuedB | smA: enter sfull ld done | smB: enter sfull ld done   (cycles from start)")
for j in range(n):
    m = [int(x) - t0 for x in t[0, j]]
    a = [int(x) - t0 for x in t[1, j]]
    bb = [int(x) - t0 for x in t[2, j]]
    print(f"{j:3d} | {m[0]:7d} {m[1]:7d} {m[2]:7d} {m[3]:7d} | {a[0]:7d} {a[1]:7d} {a[2]:7d} {a[3]:7d} | {bb[0]:7d} {bb[1]:7d} {bb[2]:7d} {bb[3]:7d}")
print("softmax A: wait-for-S / load / compute per step:",
      [(int(t[1, j, 1] - t[1, j, 0]), int(t[1, j, 2] - t[1, j, 1]), int(t[1, j, 3] - t[1, j, 2])) for j in range(4, 10)])
life = [int(x) - int(t[3, 0, 0]) for x in t[3, 0]]
steady = (int(t[0, n - 2, 0]) - int(t[0, 3, 0])) / (n - 5)
print(f"CTA life cycle (cycles from entry): set-up done {life[1]}, first S wait {t0 - int(t[3, 0, 0])}, O complete {life[2]}, "
      f"last store issued {life[3]}; steady-state step {steady:.0f} cycles; {n} steps -> "
      f"{100 * (1 - n * steady / max(1, life[3])):.1f} % of the CTA's life is prologue / epilogue / pipeline fill")
if int(t[3, 1, 0]) > 0:  # persistent kernel: item boundary of the first item (cycles from CTA entry)
    e = int(t[3, 0, 0])
    for tt, nm in ((0, "A"), (1, "B")):
        print(f"tile {nm}: O complete {int(t[3, 3 + tt, 0]) - e}, staging acquired {int(t[3, 1 + tt, 0]) - e}, staging written "
              f"{int(t[3, 1 + tt, 1]) - e}, tile barrier passed {int(t[3, 1 + tt, 2]) - e}, stores read {int(t[3, 3 + tt, 1]) - e}, "
              f"first S of the next item seen {int(t[3, 1 + tt, 3]) - e}")
    print(f"last softmax step of item 0 ends at A {int(t[1, n - 1, 3]) - e} / B {int(t[2, n - 1, 3]) - e}")
