"""Stand-alone GB/s of the LayerNorm + modulation kernel at the production shapes (algorithmic bytes: fp32 in, bf16 out),
L2 flushed between iterations by rotating over buffers larger than the 126 MB L2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ladcast_b200 import _lib
lib = _lib.load()
for rows, d, rps in ((45000, 1536, 2250), (36000, 1536, 1800), (9000, 1536, 450), (29250, 2048, 2250), (6750, 2048, 2250)):
    nb = (rows + rps - 1) // rps
    xs = [torch.randn(rows, d, device="cuda") for _ in range(3)]
    outs = [torch.empty(rows, d, device="cuda", dtype=torch.bfloat16) for _ in range(3)]
    mod = torch.randn(nb, 3 * d, device="cuda")
    def run(i):
        _lib.check(lib.lc_layernorm_modulate(_lib.PRECISION_BF16, _lib.ptr(xs[i % 3]), _lib.ptr(outs[i % 3]), rows, d, 1e-6, rps,
                                             _lib.ptr_any(mod[:, d:]), _lib.ptr(mod), 3 * d, None, None, _lib.stream()), "ln")
    for i in range(6): run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 60
    e0.record()
    for i in range(n): run(i)
    e1.record(); torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / n
    print(f"layernorm rows={rows} d={d}: {us:.1f} us  {rows * d * 6 / us / 1e3:.0f} GB/s ({os.environ.get('LADCAST_B200_LN_ROWS', 'interleave')})")
