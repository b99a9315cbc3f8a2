"""GPU parity tests of the low-level kernels through the C ABI (lc_gemm, lc_attention, scheduler steps)."""
import ctypes

import pytest
import torch

from conftest import record_measured
from ladcast_b200 import _lib

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


@pytest.fixture(scope="module")
def lib():
    return _lib.load()


@pytest.mark.parametrize("m,n,k", [(128, 256, 64), (300, 512, 256), (1800, 1536, 1536), (20, 4608, 1536),
                                   (450, 84, 1536), (900, 1008, 96), (257, 6144, 320)])
@pytest.mark.parametrize("prec", [_lib.PRECISION_F32, _lib.PRECISION_BF16], ids=["f32", "bf16"])
def test_gemm(lib, m, n, k, prec):
    g = torch.Generator("cpu").manual_seed(m * 7 + n * 3 + k)
    a = torch.randn(m, k, generator=g).cuda()
    w = (torch.randn(n, k, generator=g) / k**0.5).cuda()
    bias = torch.randn(n, generator=g).cuda()
    c = torch.full((m, n), float("nan"), device="cuda")
    if prec == _lib.PRECISION_BF16:
        a_in, w_in = a.bfloat16().contiguous(), w.bfloat16().contiguous()
        ref = torch.nn.functional.gelu(a_in.float() @ w_in.float().T + bias, approximate="tanh")
        tol = 2e-5
    else:
        a_in, w_in = a, w
        ref = torch.nn.functional.gelu(a.double() @ w.double().T + bias.double(), approximate="tanh").float()
        tol = 2e-6
    _lib.check(lib.lc_gemm(prec, _lib.ptr(a_in), _lib.ptr(w_in), _lib.ptr(bias), _lib.ptr(c), m, n, k, 1, _lib.stream()),
               "lc_gemm")
    torch.cuda.synchronize()
    assert torch.isfinite(c).all()
    assert _rel(c, ref) < tol


@pytest.mark.parametrize("b,s,heads", [(1, 128, 1), (2, 450, 2), (1, 2250, 3), (3, 200, 1), (1, 90, 2), (2, 129, 1), (1, 257, 1)])
@pytest.mark.parametrize("prec", [_lib.PRECISION_F32, _lib.PRECISION_BF16], ids=["f32", "bf16"])
def test_attention(lib, b, s, heads, prec):
    g = torch.Generator("cpu").manual_seed(b * 100 + s + heads)
    d = heads * 128
    qkv = torch.randn(b, s, 3 * d, generator=g).cuda()
    qkv[..., : 2 * d] *= 1.5  # some spread in the logits
    if prec == _lib.PRECISION_BF16:
        qkv_in = qkv.bfloat16().contiguous()
        out = torch.zeros(b, s, d, device="cuda", dtype=torch.bfloat16)
        tol = 3e-3  # measured 1.8e-3 .. 2.3e-3 (profiles/r02_measured_parity.jsonl); P and the output are bf16
    else:
        qkv_in = qkv
        out = torch.zeros(b, s, d, device="cuda")
        tol = 1e-5
    qf = qkv_in.double()
    q, k, v = [t.reshape(b, s, heads, 128).transpose(1, 2) for t in qf.chunk(3, dim=-1)]
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(b, s, d)
    _lib.check(lib.lc_attention(prec, _lib.ptr(qkv_in), _lib.ptr(out), b, s, heads, _lib.stream()), "lc_attention")
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    r = _rel(out, ref)
    if prec == _lib.PRECISION_BF16:
        record_measured(f"attention/bf16/b{b}_s{s}_h{heads}/rel_l2", r)
    assert r < tol


def test_dpmpp2m_step(lib):
    from oracle import ladcast_oracle as O

    g = torch.Generator("cpu").manual_seed(5)
    shape = (3, 84, 2, 15, 30)
    noise = torch.randn(shape, generator=g)
    fs = [torch.randn(shape, generator=g) for _ in range(6)]
    it = iter(fs)
    want = O.dpmpp2m_sample(lambda xin, cn: next(it), noise, 6)
    from ladcast_b200.pipelines.scheduler import dpmpp2m_coefficients

    x = noise.clone().cuda()
    x0p = torch.zeros_like(x)
    xin = torch.empty_like(x)
    for i in range(6):
        c = dpmpp2m_coefficients(6, i)
        fd = fs[i].cuda()
        _lib.check(lib.lc_sched_dpmpp2m_step(_lib.ptr(fd), _lib.ptr(x), _lib.ptr(x0p), _lib.ptr(xin), x.numel(),
                                             c["c_skip"], c["c_out"], c["a_x"], c["a_x0"], c["a_d"], c["c_in_next"],
                                             _lib.stream()), "sched")
    torch.cuda.synchronize()
    assert _rel(x.cpu(), want) < 1e-6


@pytest.mark.parametrize("rows,d,rps", [(37, 256, 10), (450, 1536, 450), (2250 * 3 + 5, 1536, 2250), (1000, 2048, 333)])
@pytest.mark.parametrize("prec", [_lib.PRECISION_F32, _lib.PRECISION_BF16], ids=["f32", "bf16"])
def test_layernorm_modulate(lib, rows, d, rps, prec):
    """LayerNorm (no affine) + AdaLN modulation, and the affine variant, against torch in float64."""
    g = torch.Generator("cpu").manual_seed(rows + d)
    x = (torch.randn(rows, d, generator=g) * 2 + 0.3).cuda()
    nb = (rows + rps - 1) // rps
    mod = torch.randn(nb, 3 * d, generator=g).cuda() * 0.5
    w, b = torch.randn(d, generator=g).cuda(), torch.randn(d, generator=g).cuda()
    dt = torch.float32 if prec == _lib.PRECISION_F32 else torch.bfloat16
    tol = 2e-6 if prec == _lib.PRECISION_F32 else 4e-3
    ln = torch.nn.functional.layer_norm(x.double(), (d,), eps=1e-6)
    sample = torch.arange(rows, device="cuda") // rps
    want_mod = ln * (1 + mod[sample, d : 2 * d].double()) + mod[sample, :d].double()
    out = torch.full((rows, d), float("nan"), device="cuda", dtype=dt)
    _lib.check(lib.lc_layernorm_modulate(prec, _lib.ptr(x), _lib.ptr(out), rows, d, 1e-6, rps, _lib.ptr_any(mod[:, d:]), _lib.ptr(mod),
                                         3 * d, None, None, _lib.stream()), "lc_layernorm_modulate")
    torch.cuda.synchronize()
    assert _rel(out, want_mod) < tol
    _lib.check(lib.lc_layernorm_modulate(prec, _lib.ptr(x), _lib.ptr(out), rows, d, 1e-6, 1 << 30, None, None, 0, _lib.ptr(w),
                                         _lib.ptr(b), _lib.stream()), "lc_layernorm_modulate")
    torch.cuda.synchronize()
    assert _rel(out, ln * w.double() + b.double()) < tol
