"""GPU parity of the DC-AE decoder path (sphere conv implicit GEMM, decoder handle, decode_latent_ens)."""
import os

import numpy as np
import pytest
import torch

from ladcast_b200 import _lib
from conftest import record_measured
from oracle import ladcast_oracle as O

pytestmark = pytest.mark.gpu


def _seeded(shape, seed, scale=1.0):
    return torch.randn(shape, generator=torch.Generator("cpu").manual_seed(seed)) * scale


def _rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("n,cin,H,W,cout", [(3, 64, 5, 8, 48), (2, 84, 15, 30, 1008), (5, 168, 15, 30, 96),
                                            (3, 40, 30, 60, 300), (1, 252, 120, 240, 252), (2, 100, 60, 120, 89)])
@pytest.mark.parametrize("prec", ["f32", "bf16"])
def test_sphere_conv3x3(n, cin, H, W, cout, prec):
    lib = _lib.load()
    x = _seeded((n, cin, H, W), 1 + cin)
    w = O.det_tensor("t.weight", (cout, cin, 3, 3), cin)
    b = O.det_tensor("t.bias", (cout,), cin)
    if prec == "bf16":
        xr, wr, tol = x.bfloat16().float(), w.bfloat16().float(), 1e-4
    else:
        xr, wr, tol = x, w, 1e-5
    want = torch.nn.functional.silu(O.sphere_conv(xr.double(), wr.double(), b.double()))
    out = torch.full((n, cout, H, W), float("nan"), device="cuda")
    xd, wd, bd = x.cuda(), w.cuda(), b.cuda()  # keep the device tensors alive across the call
    _lib.check(lib.lc_sphere_conv3x3(_lib.PRECISION_F32 if prec == "f32" else _lib.PRECISION_BF16, _lib.ptr(xd),
                                     _lib.ptr(wd), _lib.ptr(bd), _lib.ptr(out), n, cin, H, W, cout, 2,
                                     _lib.stream()), "lc_sphere_conv3x3")
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    assert _rel(out, want) < tol


def _ae(name, salt, precision):
    from ladcast_b200.models import AutoencoderDC

    cfg = O.dcae_config(name) if isinstance(name, str) else name
    sd = O.make_state_dict(O.dcae_decoder_param_shapes(cfg), salt)
    ae = AutoencoderDC(**cfg)
    ae.load_state_dict(sd, strict=True)
    return cfg, sd, ae.to("cuda").set_precision(precision)


def test_dcae_tiny_vs_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "dcae_tiny.npz"))
    cfg, sd, ae = _ae("tiny", int(g["salt"]), "fp32")
    out = ae.decode(_seeded((2, 84, 5, 8), 103).cuda()).sample
    torch.cuda.synchronize()
    assert out.shape == (2, 84, 40, 64)
    assert _rel(out[:, ::7], g["out_sub"]) < 1e-4
    assert np.allclose(out.double().sum(dim=(0, 2, 3)).cpu().numpy(), g["ch_sum"], rtol=1e-3, atol=5e-2)
    from ladcast_b200.pipelines.utils import decode_latent_ens

    mean, std = _seeded((84,), 105), _seeded((84,), 106).abs() + 0.5
    ens = decode_latent_ens(ae, _seeded((1, 84, 2, 5, 8), 104).cuda(), mean, std)
    assert ens.shape == (1, 84, 2, 40, 64)
    assert _rel(ens[:, ::7], g["ens_sub"]) < 1e-4


SMALL = O.dcae_config("tiny")


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 8.5e-3)])  # measured 1.4e-6 / 6.3e-3
def test_dcae_small_vs_oracle(precision, tol):
    cfg, sd, ae = _ae(SMALL, 22, precision)
    z = _seeded((3, 84, 15, 30), 500)
    taps = {}
    want = O.dcae_decode(sd, cfg, z, taps=taps)
    out = ae.decode(z.cuda()).sample
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    r = _rel(out, want)
    record_measured(f"dcae_small/{precision}/rel_l2", r)
    assert r < tol


def test_dcae_full_bf16_vs_oracle():
    """V0.1.X decoder architecture (143.2 M parameters), one frame, bf16 tensor-core path vs the fp32 oracle."""
    cfg, sd, ae = _ae("V0.1.X", 23, "bf16")
    z = _seeded((1, 84, 15, 30), 600)
    want = O.dcae_decode(sd, cfg, z)
    mean, std = _seeded((84,), 105), _seeded((84,), 106).abs() + 0.5
    out = ae.decode(z.cuda()).sample
    fused = ae.decode_fused(z.cuda(), mean, std)
    torch.cuda.synchronize()
    r = _rel(out, want)
    record_measured("dcae_full_V0.1.X/bf16/rel_l2", r)
    assert r < 1e-2  # north_star per-call budget; measured 7.9e-3 (profiles/r02_measured_parity.jsonl)
    assert _rel(fused, want * std[None, :, None, None] + mean[None, :, None, None]) < 1e-2


def test_dcae_decode_160_latents_batch_consistency():
    """BASELINE config 3: decoder-only batch decode of 160 ensemble latents -> 84x120x240 fields (V0.1.X
    architecture).  Size-independent property: a frame decoded inside the batch (any chunk / tile grouping) is
    bit-identical to the same frame decoded alone, and de-normalisation commutes with the fused epilogue."""
    cfg, sd, ae = _ae("V0.1.X", 23, "bf16")
    z = _seeded((160, 84, 15, 30), 700).cuda()
    out = ae.decode(z).sample
    torch.cuda.synchronize()
    assert out.shape == (160, 84, 120, 240) and torch.isfinite(out).all()
    for i in (0, 79, 80, 159):
        one = ae.decode(z[i : i + 1].contiguous()).sample
        assert torch.equal(one[0], out[i]), i
    mean, std = _seeded((84,), 105), _seeded((84,), 106).abs() + 0.5
    fused = ae.decode_fused(z[:4].contiguous(), mean, std)
    ref = out[:4] * std.cuda()[None, :, None, None] + mean.cuda()[None, :, None, None]
    assert torch.allclose(fused, ref, rtol=1e-5, atol=1e-5)


# ---------------------------------------------------------------------------------------------------------------
# encoder (AutoencoderDC.encode, DCAE.py:964-1000 / Encoder.forward :617-631)
# ---------------------------------------------------------------------------------------------------------------
def _ae_enc(name, salt, precision):
    from ladcast_b200.models import AutoencoderDC

    cfg = O.dcae_config(name) if isinstance(name, str) else name
    sd = O.make_state_dict(O.dcae_encoder_param_shapes(cfg), salt)
    ae = AutoencoderDC(**cfg)
    ae.load_state_dict(sd, strict=True)
    return cfg, sd, ae.to("cuda").set_precision(precision)


def test_dcae_encode_tiny_vs_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "dcae_encode_tiny.npz"))
    cfg, sd, ae = _ae_enc("tiny", int(g["salt"]), "fp32")
    x, static = _seeded((2, 84, 40, 64), 107).cuda(), _seeded((2, 5, 40, 64), 108).cuda()
    lat = ae.encode(x, static_conditioning_tensor=static).latent
    torch.cuda.synchronize()
    assert lat.shape == (2, 84, 5, 8)
    assert _rel(lat, g["latent"]) < 1e-4
    lat2 = ae.encode(torch.cat((x, static), dim=1), return_dict=False)[0]
    assert torch.equal(lat, lat2)


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 2e-2)])
def test_dcae_encode_small_vs_oracle(precision, tol):
    """Different geometry / batch than the golden, plus the fused latent normalisation."""
    cfg, sd, ae = _ae_enc("tiny", 31, precision)
    x = _seeded((3, 89, 48, 80), 211)
    want = O.dcae_encode(sd, cfg, x)
    got = ae.encode(x.cuda()).latent
    assert got.shape == (3, 84, 6, 10)
    assert _rel(got, want) < tol
    mean, std = _seeded((84,), 212), _seeded((84,), 213).abs() + 0.5
    want_n = O.normalize_latent(want.unsqueeze(2), mean, std, 0.5).squeeze(2)
    got_n = ae.encode_fused(x.cuda(), mean, std, 0.5)
    assert _rel(got_n, want_n) < tol


def test_dcae_encode_full_config_bf16():
    """V0.1.X encoder (4x4 layers, 252/504/504/1008) on one 120x240 frame against the CPU oracle."""
    cfg, sd, ae = _ae_enc("V0.1.X", 33, "bf16")
    x = _seeded((1, 89, 120, 240), 214)
    want = O.dcae_encode(sd, cfg, x)
    got = ae.encode(x.cuda()).latent
    assert got.shape == (1, 84, 15, 30)
    assert _rel(got, want) < 2e-2


def test_dcae_roundtrip_shapes_and_determinism():
    """encode -> decode through one handle holding both halves; repeated calls are bit-identical."""
    from ladcast_b200.models import AutoencoderDC

    cfg = O.dcae_config("tiny")
    sd = O.make_state_dict(O.dcae_encoder_param_shapes(cfg), 35)
    sd.update(O.make_state_dict(O.dcae_decoder_param_shapes(cfg), 36))
    ae = AutoencoderDC(**cfg)
    ae.load_state_dict(sd, strict=True)
    ae.to("cuda").set_precision("bf16")
    x = _seeded((2, 89, 40, 64), 215).cuda()
    z1 = ae.encode(x).latent
    y1 = ae.decode(z1).sample
    z2 = ae.encode(x).latent
    assert torch.equal(z1, z2)
    assert y1.shape == (2, 84, 40, 64) and torch.isfinite(y1).all()
    want = O.dcae_decode(sd, cfg, O.dcae_encode(sd, cfg, x.cpu()))
    assert _rel(y1, want) < 3e-2
