import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ladcast_b200 import _lib
lib = _lib.load()
b, s, h = 20, 2250, 12
d = h * 128
qkv = torch.randn(b, s, 3 * d, device="cuda").bfloat16()
out = torch.empty(b, s, d, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    _lib.check(lib.lc_attention(0, _lib.ptr(qkv), _lib.ptr(out), b, s, h, _lib.stream()))
torch.cuda.synchronize()
