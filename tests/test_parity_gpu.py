"""GPU parity tests added in round 2 (through the drop-in classes / the C ABI):
  * north-star tolerance #2: ensemble-mean lat-weighted RMSE and CRPS of a BF16 rollout within 1 % of the oracle's;
  * goldens from the unmodified reference: get_acc, Heun N=8, ladcast_1.6B denoiser;
  * index permutations asserted bit-exact (patchify, unpatchify, pixel shuffle / unshuffle, 5-D decode addressing,
    latent feedback);
  * the metrics kernel over ensemble sizes on both code paths (sorted registers / pairwise shared memory), strided input.
"""
import os

import numpy as np
import pytest
import torch

from conftest import record_measured
from ladcast_b200 import _lib
from oracle import ladcast_oracle as O

pytestmark = pytest.mark.gpu


def _seeded(shape, seed, scale=1.0):
    return torch.randn(shape, generator=torch.Generator("cpu").manual_seed(seed)) * scale


def _rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _denoiser(name, salt, precision):
    from ladcast_b200.models import LaDCastTransformer3DModel

    cfg = O.denoiser_config(name)
    sd = O.make_state_dict(O.denoiser_param_shapes(cfg), salt)
    m = LaDCastTransformer3DModel.from_config(cfg)
    m.load_state_dict(sd, strict=True)
    return cfg, sd, m.to("cuda").set_precision(precision)


def _autoencoder(name, precision, salt=21):
    from ladcast_b200.models import AutoencoderDC

    acfg = O.dcae_config(name)
    asd = O.make_state_dict(O.dcae_decoder_param_shapes(acfg), salt)
    ae = AutoencoderDC(**acfg)
    ae.load_state_dict(asd)
    return acfg, asd, ae.to("cuda").set_precision(precision)


# ------------------------------------------------------------------------------------------- north-star tolerance #2
def _rollout_metrics_case(den_name, den_salt, ae_name, ens, n_ar, t_out, n_steps, tag):
    """BF16 rollout (sampler -> AR feedback -> de-normalise -> decode) -> ensemble metrics on the device, against the
    fp32 oracle rollout and oracle metrics (train_AR.py:281-312, evaluate_ens_gpu.py:351-415).  Truth = a second seeded
    N(0,1) field tensor de-normalised like the forecast.  RMSE = sqrt(ens_mse) and CRPS per (channel,
    lead) must agree within 1 %."""
    from ladcast_b200.evaluate.utils import ensemble_metrics
    from ladcast_b200.pipelines import AutoRegressive2DPipeline, EDMDPMSolverMultistepScheduler
    from ladcast_b200.pipelines.utils import roll_out_latent, rollout_as_lead_major

    cfg, sd, m = _denoiser(den_name, den_salt, "bf16")
    acfg, asd, ae = _autoencoder(ae_name, "bf16")
    pipe = AutoRegressive2DPipeline(m, EDMDPMSolverMultistepScheduler())
    known = _seeded((1, 84, 1, 15, 30), 102, 0.5)
    lat_mean, lat_std = _seeded((84,), 31) * 0.1, _seeded((84,), 32).abs() + 0.5
    f_mean, f_std = _seeded((84,), 33), _seeded((84,), 34).abs() + 0.5
    members = list(range(ens))
    total = n_ar * t_out
    out = roll_out_latent(pipe, ae, known, 2018010100, ens, lat_mean, lat_std, f_mean, f_std, num_inference_steps=n_steps,
                          return_seq_len=t_out, total_lead_time_hour=6 * total, member_indices=members)
    got = rollout_as_lead_major(out)  # (ens, 84, total, 120, 240) on the host
    _, want = O.rollout(sd, cfg, asd, acfg, known, members, 2018010100, total, t_out, n_steps, lat_mean, lat_std, f_mean,
                        f_std, sampler="pipeline")
    assert got.shape == want.shape
    truth = _seeded((84, total, 120, 240), 777) * f_std[:, None, None, None] + f_mean[:, None, None, None]
    tabs = ensemble_metrics(got.cuda(), truth.cuda())
    ref = O.ensemble_metrics(want, truth)
    rmse_g, rmse_w = tabs["ens_mse"].cpu().sqrt(), ref["ens_mse"].sqrt()
    d_rmse = float(((rmse_g - rmse_w).abs() / rmse_w).max())
    d_crps = float(((tabs["crps"].cpu() - ref["crps"]).abs() / ref["crps"].abs()).max())
    d_spread = float(((tabs["crps_spread"].cpu() - ref["crps_spread"]).abs() / ref["crps_spread"].abs()).max())
    field_rel = _rel(got.mean(0), want.mean(0))
    record_measured(f"rollout_metrics/{tag}/rmse_max_rel", d_rmse)
    record_measured(f"rollout_metrics/{tag}/crps_max_rel", d_crps)
    record_measured(f"rollout_metrics/{tag}/spread_max_rel", d_spread)
    record_measured(f"rollout_metrics/{tag}/ens_mean_field_rel_l2", field_rel)
    assert torch.isfinite(got).all()
    assert d_rmse < 1e-2, f"ensemble-mean RMSE differs by {d_rmse:.3e} (> 1 %)"
    assert d_crps < 1e-2, f"CRPS differs by {d_crps:.3e} (> 1 %)"


def test_bf16_rollout_metrics_within_1pct_tiny():
    """tiny denoiser + tiny DC-AE, BF16: 3 AR steps x 20 solver steps x 8 members, T_out = 2 (6 lead steps)."""
    _rollout_metrics_case("tiny", 11, "tiny", ens=8, n_ar=3, t_out=2, n_steps=20, tag="tiny_3ar_20steps_ens8")


def test_bf16_rollout_metrics_within_1pct_375M():
    """ladcast_375M geometry, BF16: one AR step of the production shape (T_out = 4, 20 solver steps), 2 members, tiny DC-AE."""
    _rollout_metrics_case("375M", 12, "tiny", ens=2, n_ar=1, t_out=4, n_steps=20, tag="375M_1ar_20steps_ens2")


# ------------------------------------------------------------------------------------------- reference goldens
def test_get_acc_vs_reference_golden(golden_dir):
    from ladcast_b200.evaluate.utils import get_acc

    g = np.load(os.path.join(golden_dir, "acc.npz"))
    f, t, c = _seeded((84, 120, 24), 21), _seeded((84, 120, 24), 22), _seeded((84, 120, 24), 23, 0.3)
    t[82, :4] = float("nan")
    w = torch.from_numpy(O.lat_weights(120)).view(-1, 1)
    got_w = get_acc(f.cuda(), t.cuda(), c.cuda(), w.cuda())
    got_u = get_acc(f.cuda(), t.cuda(), c.cuda())
    assert np.allclose(got_w.cpu().numpy(), g["weighted"], rtol=1e-6, atol=1e-9)
    assert np.allclose(got_u.cpu().numpy(), g["unweighted"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-3), ("bf16", 5e-2)])
def test_heun8_vs_reference_golden(golden_dir, precision, tol):
    """EDM Heun sampler, N = 8 (15 denoiser calls, fp64 state): FP32 validation mode against the unmodified reference;
    the BF16 path must stay close to it (sampler errors compound over 15 calls)."""
    from ladcast_b200.pipelines import AutoRegressive2DPipeline, EDMDPMSolverMultistepScheduler
    from ladcast_b200.pipelines.utils import ensemble_AR_sampler

    g = np.load(os.path.join(golden_dir, "heun8_tiny.npz"))
    cfg, sd, m = _denoiser("tiny", 11, precision)
    pipe = AutoRegressive2DPipeline(m, EDMDPMSolverMultistepScheduler())
    s = ensemble_AR_sampler(pipe, sample_size=2, return_seq_len=1, num_inference_steps=8,
                            known_latents=_seeded((1, 84, 1, 15, 30), 102, 0.5).cuda(),
                            timestamps=torch.tensor([2018010100]), sampler_type="edm", device="cuda")
    torch.cuda.synchronize()
    r = _rel(s, g["edm_8"])
    record_measured(f"heun8/{precision}/rel_l2", r)
    assert r < tol


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-3), ("bf16", 5e-2)])
def test_heun_churn_vs_reference_golden(golden_dir, precision, tol):
    """Stochastic churn (deterministic=False, edm_sampler.py:67-76): `lc_sched_heun_churn` + the Heun kernels against the
    unmodified reference on the same seeded fp64 churn noise (drawn on the CPU in call order, moved to the device)."""
    from ladcast_b200.pipelines import AutoRegressive2DPipeline, EDMDPMSolverMultistepScheduler
    from ladcast_b200.pipelines.utils import ensemble_AR_sampler

    g = np.load(os.path.join(golden_dir, "heun_churn_tiny.npz"))
    cfg, sd, m = _denoiser("tiny", 11, precision)
    pipe = AutoRegressive2DPipeline(m, EDMDPMSolverMultistepScheduler())
    gen = torch.Generator("cpu").manual_seed(777)
    kw = dict(deterministic=False, S_churn=4.0, S_min=0.05, S_max=50.0, S_noise=1.003,
              randn_like=lambda x: torch.randn(x.shape, generator=gen, dtype=x.dtype).to(x.device))
    s = ensemble_AR_sampler(pipe, sample_size=2, return_seq_len=1, num_inference_steps=6,
                            known_latents=_seeded((1, 84, 1, 15, 30), 102, 0.5).cuda(),
                            timestamps=torch.tensor([2018010100]), sampler_type="edm", device="cuda", sampler_kwargs=kw)
    torch.cuda.synchronize()
    r = _rel(s, g["edm_churn_6"])
    record_measured(f"heun_churn/{precision}/rel_l2", r)
    assert r < tol


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_denoiser_1p6B_vs_reference_golden(golden_dir, precision):
    g = np.load(os.path.join(golden_dir, "denoiser_1p6B.npz"))
    cfg, sd, m = _denoiser("1.6B", int(g["salt"]), precision)
    x = _seeded((1, 84, 1, 15, 30), 100).cuda()
    cond = _seeded((1, 84, 1, 15, 30), 101, 0.5).cuda()
    out = m(x, torch.from_numpy(g["t"]).cuda(), cond, time_elapsed=torch.from_numpy(g["ts"]), return_dict=False)[0]
    torch.cuda.synchronize()
    r = _rel(out, g["out"])
    record_measured(f"denoiser_1p6B_golden/{precision}/rel_l2", r)
    assert r < (1e-4 if precision == "fp32" else 1e-2)


# ------------------------------------------------------------------------------------------- bit-exact permutations
def test_patchify_unpatchify_bit_exact():
    """north star: 'index/patch permutations bit-exact'.  Tokens == x.flatten(2).transpose(1, 2) (embeddings.py:56-59,
    token n = t*450 + h*30 + w) and an identity proj_out reproduces the input tensor exactly (LaDCast_3D_model.py:
    1047-1062, patch size 1) — in the FP32 mode and, for bf16-representable inputs, on the tensor-core path."""
    lib = _lib.load()
    B, C, T, H, W, Kp = 3, 84, 4, 15, 30, 96
    x = _seeded((B, C, T, H, W), 900).bfloat16().float().cuda()  # exactly representable in bf16
    want_tok = x.flatten(2).transpose(1, 2)  # (B, THW, C)
    for prec, dt in ((_lib.PRECISION_F32, torch.float32), (_lib.PRECISION_BF16, torch.bfloat16)):
        tok = torch.full((B * T * H * W, Kp), float("nan"), device="cuda", dtype=dt)
        _lib.check(lib.lc_patchify(prec, _lib.ptr(x), _lib.ptr(tok), B, C, T * H * W, Kp, _lib.stream()), "lc_patchify")
        torch.cuda.synchronize()
        assert torch.equal(tok[:, :C].float().reshape(B, T * H * W, C), want_tok)
        assert bool((tok[:, C:] == 0).all())
        eye = torch.zeros((C, Kp), device="cuda", dtype=dt)
        eye[torch.arange(C), torch.arange(C)] = 1
        out = torch.full((B, C, T, H, W), float("nan"), device="cuda")
        _lib.check(lib.lc_unpatchify_gemm(prec, _lib.ptr(tok), _lib.ptr(eye), None, _lib.ptr(out), B, T * H * W, C, Kp,
                                          _lib.stream()), "lc_unpatchify_gemm")
        torch.cuda.synchronize()
        assert torch.equal(out, x)
        # a channel permutation as weight: output channel c must be input channel perm[c], nothing else moves
        perm = torch.randperm(C, generator=torch.Generator("cpu").manual_seed(5)).cuda()
        pw = torch.zeros((C, Kp), device="cuda", dtype=dt)
        pw[torch.arange(C, device="cuda"), perm] = 1
        _lib.check(lib.lc_unpatchify_gemm(prec, _lib.ptr(tok), _lib.ptr(pw), None, _lib.ptr(out), B, T * H * W, C, Kp,
                                          _lib.stream()), "lc_unpatchify_gemm")
        torch.cuda.synchronize()
        assert torch.equal(out, x[:, perm])


@pytest.mark.parametrize("cin,cout", [(336, 168), (168, 168), (168, 84), (1008, 504)])
def test_pixel_shuffle_shortcut_bit_exact(cin, cout):
    """DCUpBlock2d tail (DCAE.py:519-536): pixel_shuffle(conv, 2) + pixel_shuffle(repeat_interleave(x, 4*cout/cin), 2)."""
    import torch.nn.functional as F

    lib = _lib.load()
    n, H, W = 2, 6, 10
    conv = _seeded((n, 4 * cout, H, W), 910).cuda()
    x = _seeded((n, cin, H, W), 911).cuda()
    want = F.pixel_shuffle(conv, 2) + F.pixel_shuffle(x.repeat_interleave(4 * cout // cin, dim=1), 2)
    out = torch.full((n, 2 * H, 2 * W, cout), float("nan"), device="cuda")
    conv_l, x_l = conv.permute(0, 2, 3, 1).contiguous(), x.permute(0, 2, 3, 1).contiguous()  # NHWC (kept alive)
    _lib.check(lib.lc_pixel_shuffle_shortcut(_lib.ptr(conv_l), _lib.ptr(x_l), _lib.ptr(out), n, H, W, cin, cout,
                                             _lib.stream()), "lc_pixel_shuffle_shortcut")
    torch.cuda.synchronize()
    assert torch.equal(out.permute(0, 3, 1, 2), want)


@pytest.mark.parametrize("cin,cout", [(84, 168), (168, 168), (168, 336)])
def test_pixel_unshuffle_shortcut_index_map(cin, cout):
    """DCDownBlock2d tail (DCAE.py:476-490).  The conv term is a pure permutation (bit-exact); the shortcut is a mean
    over 4*cin/cout channels (summation order may differ from torch's by an ulp)."""
    import torch.nn.functional as F

    lib = _lib.load()
    n, H, W = 2, 8, 12
    conv = _seeded((n, cout // 4, H, W), 920).cuda()
    x = _seeded((n, cin, H, W), 921).cuda()
    out = torch.full((n, H // 2, W // 2, cout), float("nan"), device="cuda")
    args = (n, H, W, cin, cout, _lib.stream())
    conv_l, x_l = conv.permute(0, 2, 3, 1).contiguous(), x.permute(0, 2, 3, 1).contiguous()  # NHWC (kept alive)
    zeros_l = torch.zeros_like(x_l)
    _lib.check(lib.lc_pixel_unshuffle_shortcut(_lib.ptr(conv_l), _lib.ptr(zeros_l), _lib.ptr(out), *args), "unshuffle")
    torch.cuda.synchronize()
    assert torch.equal(out.permute(0, 3, 1, 2), F.pixel_unshuffle(conv, 2))
    _lib.check(lib.lc_pixel_unshuffle_shortcut(_lib.ptr(conv_l), _lib.ptr(x_l), _lib.ptr(out), *args), "unshuffle")
    torch.cuda.synchronize()
    g = 4 * cin // cout
    want = F.pixel_unshuffle(conv, 2) + F.pixel_unshuffle(x, 2).unflatten(1, (-1, g)).mean(dim=2)
    assert torch.allclose(out.permute(0, 3, 1, 2), want, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_decode_5d_in_place_equals_permuted_4d(precision):
    """decode_latent_ens through the 5-D entry point (latents read in place, fields written to (B, C, T, H, W), latent
    de-normalisation in the first kernel, extract_first) is bit-identical to permute -> decode(4-D) -> permute."""
    from ladcast_b200.pipelines.utils import decode_latent_ens

    acfg, asd, ae = _autoencoder("tiny", precision)
    lat = _seeded((3, 84, 4, 5, 8), 930).cuda()
    mean, std = _seeded((84,), 105).cuda(), (_seeded((84,), 106).abs() + 0.5).cuda()
    lm, ls = (_seeded((84,), 31) * 0.1).cuda(), (_seeded((84,), 32).abs() + 0.5).cuda()
    for take in (4, 3):
        got = decode_latent_ens(ae, lat, mean, std, extract_first=take)
        z = lat[:, :, :take].permute(0, 2, 1, 3, 4).reshape(3 * take, 84, 5, 8).contiguous()
        want = ae.decode_fused(z, mean, std).reshape(3, take, 84, 40, 64).permute(0, 2, 1, 3, 4)
        assert got.shape == (3, 84, take, 40, 64) and torch.equal(got, want)
    phys = (lat / 0.5) * ls[None, :, None, None, None] + lm[None, :, None, None, None]
    got = ae.decode_ens_fused(lat, mean, std, latent_mean=lm, latent_std=ls, target_std=0.5)
    assert torch.equal(got, decode_latent_ens(ae, phys, mean, std))
    # frames beyond one native call (MAX_FRAMES_PER_CALL) are chunked with a running frame offset
    ae.MAX_FRAMES_PER_CALL, keep = 5, ae.MAX_FRAMES_PER_CALL
    try:
        ae._reserved = None
        assert torch.equal(ae.decode_ens_fused(lat, mean, std, latent_mean=lm, latent_std=ls), got)
    finally:
        ae.MAX_FRAMES_PER_CALL = keep
        ae._reserved = None


def test_latent_feedback_bit_exact():
    """lc_latent_feedback == samples[:, :, -T_in:] and (samples / 0.5) * std + mean (pipelines/utils.py:560-577)."""
    lib = _lib.load()
    B, C, T, h, w, t_in = 3, 84, 4, 15, 30, 2
    s = _seeded((B, C, T, h, w), 940).cuda()
    lm, ls = (_seeded((C,), 31) * 0.1).cuda(), (_seeded((C,), 32).abs() + 0.5).cuda()
    known = torch.full((B, C, t_in, h, w), float("nan"), device="cuda")
    phys = torch.full_like(s, float("nan"))
    _lib.check(lib.lc_latent_feedback(_lib.ptr(s), _lib.ptr(known), _lib.ptr(phys), _lib.ptr(lm), _lib.ptr(ls), 0.5, B, C, T,
                                      t_in, h * w, _lib.stream()), "lc_latent_feedback")
    torch.cuda.synchronize()
    assert torch.equal(known, s[:, :, -t_in:])
    assert torch.equal(phys, (s / 0.5) * ls[None, :, None, None, None] + lm[None, :, None, None, None])


# ------------------------------------------------------------------------------------------- metrics kernel paths
@pytest.mark.parametrize("M", [3, 8, 17, 33, 64, 65, 100])
def test_metrics_member_counts_and_strides(M):
    """Every padded network size (<= 64 members: sorted in registers) and the pairwise shared-memory path (> 64),
    from a contiguous tensor and in place from a member-strided slice of a larger one."""
    from ladcast_b200.evaluate.utils import ensemble_metrics

    f = _seeded((M, 84, 1, 120, 12), 40 + M)
    t = _seeded((84, 1, 120, 12), 41 + M)
    t[82, 0, :3] = float("nan")
    want = O.ensemble_metrics(f, t)
    got = ensemble_metrics(f.cuda(), t.cuda())
    big = torch.full((M, 2, 84, 1, 120, 12), float("nan"), device="cuda")
    big[:, 1] = f.cuda()
    got_s = ensemble_metrics(big[:, 1], t.cuda())  # member stride = 2 members' worth of data, no copy
    for k in want:
        assert np.allclose(got[k].cpu().numpy(), want[k].numpy(), rtol=1e-5, atol=1e-8, equal_nan=True), (k, M)
        assert torch.equal(got[k], got_s[k]) or np.allclose(got[k].cpu().numpy(), got_s[k].cpu().numpy(), rtol=1e-12), (k, M)


def test_metrics_nan_member_propagates():
    """A NaN forecast value poisons that pixel's mean / skill / spread like torch.sort + sum do in the reference."""
    from ladcast_b200.evaluate.utils import get_crps, pointwise_crps_spread

    f = _seeded((6, 4, 16, 8), 60)
    f[2, 1, 3, 4] = float("nan")
    sp = pointwise_crps_spread(f.cuda(), ensemble_dim=0).cpu()
    want = O.crps_spread_pointwise(f)
    assert torch.isnan(sp[1, 3, 4]) and torch.isnan(want[1, 3, 4])
    mask = ~torch.isnan(want)
    assert torch.allclose(sp[mask], want[mask], rtol=1e-5, atol=1e-6)
    assert torch.isnan(get_crps(f.cuda(), torch.zeros(1, 4, 16, 8).cuda(), 0).cpu()[1, 3, 4])


# ------------------------------------------------------------------------------------------- roll_out_serial options
def _tiny_rollout_setup(precision="fp32"):
    from ladcast_b200.models import AutoencoderDC
    from ladcast_b200.pipelines import AutoRegressive2DPipeline, EDMDPMSolverMultistepScheduler

    cfg, sd, m = _denoiser("tiny", 11, precision)
    acfg = O.dcae_config("tiny")
    asd = O.make_state_dict(O.dcae_decoder_param_shapes(acfg), 21)
    asd.update(O.make_state_dict(O.dcae_encoder_param_shapes(acfg), 23))
    ae = AutoencoderDC(**acfg)
    ae.load_state_dict(asd)
    ae.to("cuda").set_precision(precision)
    pipe = AutoRegressive2DPipeline(m, EDMDPMSolverMultistepScheduler())
    stats = (_seeded((84,), 31) * 0.1, _seeded((84,), 32).abs() + 0.5, _seeded((84,), 33), _seeded((84,), 34).abs() + 0.5)
    return (cfg, sd, acfg, asd), pipe, ae, stats


def test_roll_out_serial_options():
    """The reference's roll_out_serial switches (pipelines/utils.py): T_out not dividing the lead count (pred_selection,
    :536-537), return_ensemble_mean (:608-630), noise_level (:518-528), several init times (:445), return_tensor layout
    with the t = 0 slot (:466-497) — FP32 validation mode, tiny models, against the oracle rollout."""
    from ladcast_b200.pipelines.utils import encode_initial_condition, roll_out_latent, roll_out_serial, rollout_as_lead_major

    (cfg, sd, acfg, asd), pipe, ae, (lm, ls, fm, fs) = _tiny_rollout_setup()
    fields = _seeded((2, 84, 1, 120, 240), 501)  # two init times, (n_init, C, T_in, H, W)
    static = _seeded((5, 120, 240), 502)
    stamps = [2018010100, 2018063018]
    kw = dict(num_inference_steps=3, return_seq_len=2, total_lead_time_hour=18)  # 3 lead steps, T_out = 2 -> blocks 2 + 1
    out, known = roll_out_serial(pipe, ae, fields.cuda(), static.cuda(), stamps, 2, lm, ls, fm, fs, **kw)
    assert out.shape == (2, 2, 2, 84, 2, 120, 240) and known.shape == (2, 84, 1, 15, 30)
    for i in range(2):
        z = O.dcae_encode(asd, acfg, fields[i].permute(1, 0, 2, 3), static.unsqueeze(0))
        kn = O.normalize_latent(z.permute(1, 0, 2, 3).unsqueeze(0), lm, ls, 0.5)
        _, want = O.rollout(sd, cfg, asd, acfg, kn, [0, 1], stamps[i], 3, 2, 3, lm, ls, fm, fs, sampler="pipeline")
        got = rollout_as_lead_major(out[i], 3)
        assert got.shape == want.shape == (2, 84, 3, 120, 240)
        assert _rel(got, want) < 2e-3
        assert torch.isnan(out[i][-1][:, :, 1]).all()  # the surplus frame of the last block is never decoded
    # ensemble mean only
    mean_out, _ = roll_out_serial(pipe, ae, fields[0].cuda(), static.cuda(), stamps[0], 2, lm, ls, fm, fs,
                                  return_ensemble_mean=True, **kw)
    assert mean_out.shape == (2, 1, 84, 2, 120, 240)
    assert torch.allclose(rollout_as_lead_major(mean_out, 3)[0], rollout_as_lead_major(out[0], 3).mean(0), rtol=1e-5, atol=1e-5)
    # reference return_tensor layout: (n_init, return_size, C, lead + 1, H, W), slot 0 = raw input field
    raw = _seeded((2, 84, 120, 240), 503)
    ref_layout = roll_out_serial(pipe, ae, fields.cuda(), static.cuda(), stamps, 2, lm, ls, fm, fs, reference_layout=True,
                                 raw_fields=raw, **kw)
    assert ref_layout.shape == (2, 2, 84, 4, 120, 240)
    assert torch.equal(ref_layout[1, 0, :, 0], raw[1]) and torch.equal(ref_layout[:, :, :, 1:], torch.stack(
        [rollout_as_lead_major(out[i], 3) for i in range(2)]))
    # latent perturbation: same as rolling out from the perturbed latents
    g = torch.Generator("cpu").manual_seed(77)
    pert, kn0 = roll_out_serial(pipe, ae, fields[0].cuda(), static.cuda(), stamps[0], 2, lm, ls, fm, fs, noise_level=0.1,
                                generator=g, return_latent=True, **kw)
    noise = torch.randn(kn0.shape, generator=torch.Generator("cpu").manual_seed(77))
    manual = roll_out_latent(pipe, ae, kn0 + (noise * 0.1 * ls.reshape(1, -1, 1, 1, 1)).cuda(), stamps[0], 2, lm, ls, fm, fs,
                             return_latent=True, **kw)
    assert torch.equal(pert[0], manual[0]) and _rel(pert[0], roll_out_latent(pipe, ae, kn0, stamps[0], 2, lm, ls, fm, fs,
                                                                               return_latent=True, **kw)[0]) > 1e-3
