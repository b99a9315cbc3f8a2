"""Micro-benchmarks of the hot kernels at production shapes (run on the B200 box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ladcast_b200 import _lib

lib = _lib.load()


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def bench_gemm(m, n, k):
    a = torch.randn(m, k, device="cuda").bfloat16()
    w = torch.randn(n, k, device="cuda").bfloat16()
    bias = torch.randn(n, device="cuda")
    c = torch.empty(m, n, device="cuda")
    t = timeit(lambda: _lib.check(lib.lc_gemm(0, _lib.ptr(a), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(c), m, n, k, 0, _lib.stream())))
    t2 = timeit(lambda: torch.matmul(a, w.T))
    print(f"gemm M={m} N={n} K={k}: {t*1e3:.3f} ms  {2*m*n*k/t/1e12:.1f} TF/s   (cuBLAS bf16 out: {t2*1e3:.3f} ms {2*m*n*k/t2/1e12:.1f} TF/s)")


def bench_attn(b, s, h):
    d = h * 128
    qkv = torch.randn(b, s, 3 * d, device="cuda").bfloat16()
    out = torch.empty(b, s, d, device="cuda", dtype=torch.bfloat16)
    t = timeit(lambda: _lib.check(lib.lc_attention(0, _lib.ptr(qkv), _lib.ptr(out), b, s, h, _lib.stream())))
    q, k, v = [x.reshape(b, s, h, 128).transpose(1, 2) for x in qkv.chunk(3, dim=-1)]
    t2 = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v))
    fl = 4 * b * h * s * s * 128
    print(f"attn B={b} S={s} H={h}: {t*1e3:.3f} ms {fl/t/1e12:.1f} TF/s   (torch SDPA: {t2*1e3:.3f} ms {fl/t2/1e12:.1f} TF/s)")


def bench_denoiser(name, B, T_out=4):
    from oracle import ladcast_oracle as O
    from ladcast_b200.models import LaDCastTransformer3DModel
    cfg = O.denoiser_config(name)
    m = LaDCastTransformer3DModel.from_config(cfg).to("cuda")
    x = torch.randn(B, 84, T_out, 15, 30, device="cuda")
    cond = torch.randn(B, 84, 1, 15, 30, device="cuda")
    t = torch.full((B,), 0.5, device="cuda")
    ts = torch.tensor([2018010100])
    with m.cached_conditioning(cond, ts, t_out=T_out):
        dt = timeit(lambda: m(x, t, cond, time_elapsed=ts), iters=5, warm=2)
    d = cfg["num_attention_heads"] * 128
    N, Nc = 450 * (T_out + 1), 450
    nb = cfg["num_layers"] + cfg["num_single_layers"]
    fl = B * (nb * (24 * d * d * N + 4 * N * N * d) + cfg["num_refiner_layers"] * (22 * d * d * Nc + 4 * Nc * Nc * d))
    print(f"denoiser {name} B={B} T_out={T_out}: {dt*1e3:.2f} ms/call  {fl/dt/1e12:.1f} TF/s (alg.)")


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    for shp in [(36000, 4608, 1536), (36000, 6144, 1536), (36000, 1536, 6144), (36000, 1536, 7680), (9000, 4608, 1536),
                (20, 58368, 1536), (36000, 84, 1536), (36000, 1536, 96)]:
        bench_gemm(*shp)
    bench_attn(20, 2250, 12)
    bench_attn(20, 450, 12)
    bench_denoiser("375M", 20)
    bench_denoiser("375M", 4)
    bench_denoiser("1.6B", 13)
