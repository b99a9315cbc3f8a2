"""GPU parity of the denoiser / samplers (through the drop-in classes -> C ABI) against the CPU oracle and the
golden vectors produced by the unmodified reference.  Tolerances are BASELINE.json's: per denoiser call
rel-L2 <= 1e-2 in bf16 and <= 1e-4 in the FP32 validation mode."""
import os

import numpy as np
import pytest
import torch

from conftest import record_measured
from oracle import ladcast_oracle as O

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-4, "bf16": 1e-2}


def _seeded(shape, seed, scale=1.0):
    return torch.randn(shape, generator=torch.Generator("cpu").manual_seed(seed)) * scale


def _rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _model(name, salt, precision):
    from ladcast_b200.models import LaDCastTransformer3DModel

    cfg = O.denoiser_config(name)
    sd = O.make_state_dict(O.denoiser_param_shapes(cfg), salt)
    m = LaDCastTransformer3DModel.from_config(cfg)
    m.load_state_dict(sd, strict=True)
    return cfg, sd, m.to("cuda").set_precision(precision)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_denoiser_tiny_vs_golden(golden_dir, precision):
    g = np.load(os.path.join(golden_dir, "denoiser_tiny.npz"))
    cfg, sd, m = _model("tiny", int(g["salt"]), precision)
    B, T_out = int(g["B"]), int(g["T_out"])
    x = _seeded((B, 84, T_out, 15, 30), 100).cuda()
    cond = _seeded((B, 84, 1, 15, 30), 101, 0.5).cuda()
    out = m(x, torch.from_numpy(g["t"]).cuda(), cond, time_elapsed=torch.from_numpy(g["ts"]), return_dict=False)[0]
    torch.cuda.synchronize()
    assert out.shape == x.shape and torch.isfinite(out).all()
    r = _rel(out, g["out"])
    record_measured(f"denoiser_tiny_golden/{precision}/rel_l2", r)
    assert r < TOL[precision]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_denoiser_tiny_t4_taps(precision):
    """T_in=1, T_out=4 (S = 2250, the production sequence) with internal taps to localise any divergence."""
    cfg, sd, m = _model("tiny", 13, precision)
    B = 2
    x = _seeded((B, 84, 4, 15, 30), 200).cuda()
    cond = _seeded((B, 84, 1, 15, 30), 201, 0.5).cuda()
    t = torch.tensor([1.0955, -1.2])
    ts = torch.tensor([2018010100, 2019063012])
    taps = {}
    want = O.denoiser_forward(sd, cfg, x.cpu(), t, cond.cpu(), ts, taps=taps)
    out = m(x, t.cuda(), cond, time_elapsed=ts).sample
    torch.cuda.synchronize()
    d = 256
    got_temb = m.debug_read("temb", B * d).reshape(B, d)
    assert _rel(got_temb, taps["temb"]) < TOL[precision], "temb"
    got_e = m.debug_read("e", B * 450 * d).reshape(B, 450, d)
    got_h = m.debug_read("h", B * 1800 * d).reshape(B, 1800, d)
    assert _rel(got_h, taps["single0.h"]) < TOL[precision] * 2, "h after last block"
    assert torch.isfinite(got_e).all()
    assert _rel(out, want) < TOL[precision]


def test_denoiser_batch_independence():
    """Member sharding must not perturb results: a member computed alone == inside a batch (bit-exact kernels)."""
    cfg, sd, m = _model("tiny", 13, "bf16")
    x = _seeded((3, 84, 1, 15, 30), 300).cuda()
    cond = _seeded((1, 84, 1, 15, 30), 301, 0.5).cuda().expand(3, -1, -1, -1, -1).contiguous()
    t = torch.tensor([0.3]).cuda()
    ts = torch.tensor([2018010100])
    full = m(x, t, cond, time_elapsed=ts).sample
    one = m(x[1:2].contiguous(), t, cond[1:2].contiguous(), time_elapsed=ts).sample
    torch.cuda.synchronize()
    assert torch.equal(full[1:2], one)


@pytest.mark.parametrize("key,sampler,n", [("pipeline_5", "pipeline", 5), ("pipeline_16", "pipeline", 16), ("edm_4", "edm", 4)])
def test_samplers_tiny_vs_golden(golden_dir, key, sampler, n):
    from ladcast_b200.pipelines import AutoRegressive2DPipeline, EDMDPMSolverMultistepScheduler
    from ladcast_b200.pipelines.utils import ensemble_AR_sampler

    g = np.load(os.path.join(golden_dir, "samplers_tiny.npz"))
    cfg, sd, m = _model("tiny", 11, "fp32")
    pipe = AutoRegressive2DPipeline(m, EDMDPMSolverMultistepScheduler())
    known = _seeded((1, 84, 1, 15, 30), 102, 0.5).cuda()
    s = ensemble_AR_sampler(pipe, sample_size=2, return_seq_len=1, num_inference_steps=n, known_latents=known,
                            timestamps=torch.tensor([2018010100]), sampler_type=sampler, device="cuda")
    torch.cuda.synchronize()
    assert _rel(s, g[key]) < 1e-3, key


def test_denoiser_375M_bf16_vs_oracle():
    """Full 375M geometry (d=1536, 12 heads, 2+4+1 blocks), production sequence, bf16 tensor-core path."""
    cfg, sd, m = _model("375M", 12, "bf16")
    B = 2
    x = _seeded((B, 84, 4, 15, 30), 400).cuda()
    cond = _seeded((B, 84, 1, 15, 30), 401, 0.5).cuda()
    t = torch.tensor([0.7159, -0.4434])
    ts = torch.tensor([2018010100])
    want = O.denoiser_forward(sd, cfg, x.cpu(), t, cond.cpu(), ts)
    out = m(x, t.cuda(), cond, time_elapsed=ts).sample
    torch.cuda.synchronize()
    r = _rel(out, want)
    record_measured("denoiser_375M/bf16/rel_l2", r)
    assert r < 1e-2


def test_denoiser_1p6B_bf16_vs_oracle():
    """ladcast_1.6B geometry (d=2048, 16 heads, 5+10+3 blocks), one member, production sequence."""
    cfg, sd, m = _model("1.6B", 14, "bf16")
    x = _seeded((1, 84, 4, 15, 30), 410).cuda()
    cond = _seeded((1, 84, 1, 15, 30), 411, 0.5).cuda()
    t = torch.tensor([0.2306])
    ts = torch.tensor([2018070112])
    want = O.denoiser_forward(sd, cfg, x.cpu(), t, cond.cpu(), ts)
    out = m(x, t.cuda(), cond, time_elapsed=ts).sample
    torch.cuda.synchronize()
    r = _rel(out, want)
    record_measured("denoiser_1p6B_T4/bf16/rel_l2", r)
    assert r < 1e-2


def test_rollout_two_ar_steps_vs_oracle():
    """roll_out_latent (sampler -> feed back -> de-normalise -> decode, two AR steps, T_out=2) vs the oracle rollout,
    in the FP32 validation mode so that the AR feedback does not amplify bf16 noise."""
    from ladcast_b200.models import AutoencoderDC
    from ladcast_b200.pipelines import AutoRegressive2DPipeline, EDMDPMSolverMultistepScheduler
    from ladcast_b200.pipelines.utils import roll_out_latent, rollout_as_lead_major

    cfg, sd, m = _model("tiny", 11, "fp32")
    acfg = O.dcae_config("tiny")
    asd = O.make_state_dict(O.dcae_decoder_param_shapes(acfg), 21)
    ae = AutoencoderDC(**acfg)
    ae.load_state_dict(asd)
    ae.to("cuda").set_precision("fp32")
    pipe = AutoRegressive2DPipeline(m, EDMDPMSolverMultistepScheduler())
    known = _seeded((1, 84, 1, 15, 30), 102, 0.5)
    lat_mean, lat_std = _seeded((84,), 31) * 0.1, _seeded((84,), 32).abs() + 0.5
    f_mean, f_std = _seeded((84,), 33), _seeded((84,), 34).abs() + 0.5
    members = [3, 4]  # global member ids (a shard of a larger ensemble)
    out = roll_out_latent(pipe, ae, known, 2018123118, 2, lat_mean, lat_std, f_mean, f_std, num_inference_steps=4,
                          return_seq_len=2, total_lead_time_hour=24, member_indices=members)
    got = rollout_as_lead_major(out)
    lat_want, fld_want = O.rollout(sd, cfg, asd, acfg, known, members, 2018123118, 4, 2, 4, lat_mean, lat_std, f_mean,
                                   f_std, sampler="pipeline")
    assert got.shape == fld_want.shape == (2, 84, 4, 120, 240)
    assert _rel(got, fld_want) < 2e-3


def test_roll_out_serial_from_fields_vs_oracle():
    """Full a1 path from standardised fields: encode (+static) -> normalise -> one AR step (T_out=2) -> decode, FP32
    validation mode, against oracle encode + normalise + rollout."""
    from ladcast_b200.models import AutoencoderDC
    from ladcast_b200.pipelines import AutoRegressive2DPipeline, EDMDPMSolverMultistepScheduler
    from ladcast_b200.pipelines.utils import roll_out_serial, rollout_as_lead_major

    cfg, sd, m = _model("tiny", 11, "fp32")
    acfg = O.dcae_config("tiny")
    asd = O.make_state_dict(O.dcae_decoder_param_shapes(acfg), 21)
    asd.update(O.make_state_dict(O.dcae_encoder_param_shapes(acfg), 23))
    ae = AutoencoderDC(**acfg)
    ae.load_state_dict(asd)
    ae.to("cuda").set_precision("fp32")
    pipe = AutoRegressive2DPipeline(m, EDMDPMSolverMultistepScheduler())
    fields = _seeded((84, 1, 120, 240), 501)  # (C, T_in, H, W), standardised units
    static = _seeded((5, 120, 240), 502)
    lat_mean, lat_std = _seeded((84,), 31) * 0.1, _seeded((84,), 32).abs() + 0.5
    f_mean, f_std = _seeded((84,), 33), _seeded((84,), 34).abs() + 0.5
    out, known = roll_out_serial(pipe, ae, fields.cuda(), static.cuda(), 2018010100, 2, lat_mean, lat_std, f_mean, f_std,
                                 num_inference_steps=4, return_seq_len=2, total_lead_time_hour=12)
    z = O.dcae_encode(asd, acfg, fields.permute(1, 0, 2, 3), static.unsqueeze(0))
    known_want = O.normalize_latent(z.permute(1, 0, 2, 3).unsqueeze(0), lat_mean, lat_std, 0.5)
    assert known.shape == known_want.shape == (1, 84, 1, 15, 30)
    assert _rel(known, known_want) < 1e-4
    _, fld_want = O.rollout(sd, cfg, asd, acfg, known_want, [0, 1], 2018010100, 2, 2, 4, lat_mean, lat_std, f_mean, f_std,
                            sampler="pipeline")
    got = rollout_as_lead_major(out)
    assert got.shape == fld_want.shape == (2, 84, 2, 120, 240)
    assert _rel(got, fld_want) < 2e-3


def test_fused_qk_epilogue_variant(golden_dir, monkeypatch):
    """LADCAST_B200_FUSE_QK=1 moves RMSNorm(q,k) + RoPE into the qkv GEMM epilogue (packed half2 cos/sin table):
    same golden, same tolerance, and within bf16 noise of the separate-kernel path."""
    g = np.load(os.path.join(golden_dir, "denoiser_tiny.npz"))
    B, T_out = int(g["B"]), int(g["T_out"])
    x = _seeded((B, 84, T_out, 15, 30), 100).cuda()
    cond = _seeded((B, 84, 1, 15, 30), 101, 0.5).cuda()
    outs = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("LADCAST_B200_FUSE_QK", flag)
        cfg, sd, m = _model("tiny", int(g["salt"]), "bf16")
        outs[flag] = m(x, torch.from_numpy(g["t"]).cuda(), cond, time_elapsed=torch.from_numpy(g["ts"]),
                       return_dict=False)[0].clone()
        torch.cuda.synchronize()
        assert _rel(outs[flag], g["out"]) < TOL["bf16"]
    assert _rel(outs["1"], outs["0"]) < 5e-3


def test_merged_stream_launches_are_bit_identical(golden_dir, monkeypatch):
    """Single-stream blocks run pred and cond tokens in one launch per projection (default) or in two
    (LADCAST_B200_MERGE_STREAMS=0): every output element sees the same K-ordered accumulation, so the results agree
    bit for bit."""
    g = np.load(os.path.join(golden_dir, "denoiser_tiny.npz"))
    B, T_out = int(g["B"]), int(g["T_out"])
    x = _seeded((B, 84, T_out, 15, 30), 100).cuda()
    cond = _seeded((B, 84, 1, 15, 30), 101, 0.5).cuda()
    outs = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("LADCAST_B200_MERGE_STREAMS", flag)
        cfg, sd, m = _model("tiny", int(g["salt"]), "bf16")
        outs[flag] = m(x, torch.from_numpy(g["t"]).cuda(), cond, time_elapsed=torch.from_numpy(g["ts"]),
                       return_dict=False)[0].clone()
        torch.cuda.synchronize()
    assert torch.equal(outs["1"], outs["0"])
