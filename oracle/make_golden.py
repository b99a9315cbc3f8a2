"""TEST INFRASTRUCTURE — golden-vector generator.  Runs the UNMODIFIED reference (/root/reference/ladcast) on
CPU fp32 with oracle/shim standing in for diffusers/xarray, on deterministic weights/inputs that
oracle/ladcast_oracle.py can rebuild from (key, shape) alone, and writes small fixtures to tests/golden/.
Run in the builder container only:  python oracle/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

from diffusers import EDMDPMSolverMultistepScheduler  # noqa: E402  (shim)
from ladcast.evaluate.utils import (  # noqa: E402
    get_acc,
    get_normalized_lat_weights_based_on_cos,
    pointwise_crps_skill,
    pointwise_crps_spread,
)
from ladcast.models.DCAE import AutoencoderDC  # noqa: E402
from ladcast.models.embeddings import LaDCastRotaryPosEmbed_from_grid, get_year_sincos_embedding  # noqa: E402
from ladcast.models.LaDCast_3D_model import LaDCastTransformer3DModel  # noqa: E402
from ladcast.models.sphere_conv import SphereConv2d  # noqa: E402
from ladcast.pipelines.pipeline_AR import AutoRegressive2DPipeline  # noqa: E402
from ladcast.pipelines.utils import decode_latent_ens, ensemble_AR_sampler  # noqa: E402

from oracle import ladcast_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
torch.set_grad_enabled(False)


def seeded(shape, seed, scale=1.0):
    return torch.randn(shape, generator=torch.Generator("cpu").manual_seed(seed)) * scale


def summary(x: torch.Tensor):
    """Size-independent fingerprints of a big tensor: per-channel (dim 1) sum and abs-sum in float64."""
    xd = x.double()
    dims = [d for d in range(x.ndim) if d != 1]
    return xd.sum(dim=dims).numpy(), xd.abs().sum(dim=dims).numpy()


def build_denoiser(name, salt):
    cfg = O.denoiser_config(name)
    m = LaDCastTransformer3DModel.from_config(cfg).eval()
    sd = O.make_state_dict(O.denoiser_param_shapes(cfg), salt)
    m.load_state_dict(sd, strict=True)  # strict: pins key names + shapes
    return cfg, m


def golden_denoiser():
    for name, salt, B, T_out in (("tiny", 11, 2, 2), ("375M", 12, 1, 1)):
        cfg, m = build_denoiser(name, salt)
        x = seeded((B, 84, T_out, 15, 30), 100)
        cond = seeded((B, 84, 1, 15, 30), 101, 0.5)
        t = torch.tensor([0.8, -0.4][:B])
        ts = torch.tensor([2018022906 if False else 2020022906])  # leap-year date
        out = m(x, t, cond, time_elapsed=ts, return_dict=False)[0]
        s, a = summary(out)
        np.savez(os.path.join(OUT, f"denoiser_{name}.npz"), salt=salt, B=B, T_out=T_out, t=t.numpy(), ts=ts.numpy(),
                 out=out.numpy().astype(np.float32), ch_sum=s, ch_abs=a)
        print("denoiser", name, out.shape, float(out.abs().mean()))


def golden_samplers():
    cfg, m = build_denoiser("tiny", 11)
    sched = EDMDPMSolverMultistepScheduler()
    pipe = AutoRegressive2DPipeline(m, sched)
    known = seeded((1, 84, 1, 15, 30), 102, 0.5)
    ts = torch.tensor([2018010100])
    res = {}
    for sampler, n in (("pipeline", 5), ("pipeline", 16), ("edm", 4)):
        s = ensemble_AR_sampler(pipe, sample_size=2, return_seq_len=1, num_inference_steps=n, known_latents=known,
                                timestamps=ts, sampler_type=sampler, device="cpu")
        res[f"{sampler}_{n}"] = s.numpy().astype(np.float32)
        print("sampler", sampler, n, float(s.abs().mean()))
    sched.set_timesteps(20)
    np.savez(os.path.join(OUT, "samplers_tiny.npz"), sigmas20=sched.sigmas.numpy(), timesteps20=sched.timesteps.numpy(),
             **res)


def golden_dcae():
    cfg = O.dcae_config("tiny")
    full = dict(cfg, encoder_block_types=cfg["decoder_block_types"],
                encoder_block_out_channels=cfg["decoder_block_out_channels"],
                encoder_layers_per_block=cfg["decoder_layers_per_block"],
                encoder_qkv_multiscales=cfg["decoder_qkv_multiscales"],
                upsample_block_type="pixel_shuffle", downsample_block_type="pixel_unshuffle")
    ae = AutoencoderDC.from_config(full).eval()
    sd = O.make_state_dict(O.dcae_decoder_param_shapes(cfg), 21)
    missing = ae.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys and all(k.startswith("encoder.") for k in missing.missing_keys)
    z = seeded((2, 84, 5, 8), 103)
    out = ae.decode(z).sample
    lat = seeded((1, 84, 2, 5, 8), 104)
    mean, std = seeded((84,), 105), seeded((84,), 106).abs() + 0.5
    ens = decode_latent_ens(ae, lat, mean, std)
    s, a = summary(out)
    np.savez(os.path.join(OUT, "dcae_tiny.npz"), salt=21, out_sub=out[:, ::7].numpy().astype(np.float32), ch_sum=s,
             ch_abs=a, ens_sub=ens[:, ::7].numpy().astype(np.float32))
    print("dcae", out.shape, float(out.abs().mean()))


def golden_dcae_encode():
    cfg = O.dcae_config("tiny")
    full = dict(cfg, upsample_block_type="pixel_shuffle", downsample_block_type="pixel_unshuffle")
    ae = AutoencoderDC.from_config(full).eval()
    sd = O.make_state_dict(O.dcae_encoder_param_shapes(cfg), 23)
    missing = ae.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys and all(k.startswith("decoder.") for k in missing.missing_keys)
    x = seeded((2, 84, 40, 64), 107)
    static = seeded((2, 5, 40, 64), 108)
    lat = ae.encode(x, static_conditioning_tensor=static).latent
    lat_cat = ae.encode(torch.cat((x, static), dim=1)).latent
    assert torch.equal(lat, lat_cat)
    np.savez(os.path.join(OUT, "dcae_encode_tiny.npz"), salt=23, latent=lat.numpy().astype(np.float32))
    print("dcae encode", lat.shape, float(lat.abs().mean()))


def golden_transforms():
    """Reference dataloader/utils.py transforms + precompute_mean_std on the shipped normalisation file."""
    import json

    from ladcast.dataloader.utils import get_inv_transform_3D, get_transform_3D, precompute_mean_std

    var_list = ["geopotential", "specific_humidity", "temperature", "u_component_of_wind", "v_component_of_wind",
                "vertical_velocity", "10m_u_component_of_wind", "10m_v_component_of_wind", "2m_temperature",
                "mean_sea_level_pressure", "sea_surface_temperature", "total_precipitation_6hr"]
    with open("/root/reference/ladcast/static/ERA5_normal_1979_2017.json") as f:
        nd = json.load(f)
    mean, std = precompute_mean_std(nd, var_list)
    # a small synthetic dict with the same structure (what the CPU test rebuilds without the reference tree)
    syn = {"a": {"mean": {"50": 1.0, "100": 2.0, "1000": -3.5}, "std": {"50": 0.5, "100": 4.0, "1000": 2.0}},
           "b": {"mean": 7.25, "std": 0.125}}
    sm, ss = precompute_mean_std(syn, ["b", "a"])
    x = seeded((4, 3, 5, 6), 120)
    args = {"mean": sm.tolist(), "std": ss.tolist(), "target_std": 0.5}
    y = get_transform_3D("normalize", args)(x)
    z = get_inv_transform_3D("normalize", args)(y)
    np.savez(os.path.join(OUT, "transforms.npz"), era5_mean=mean.numpy(), era5_std=std.numpy(), syn_mean=sm.numpy(),
             syn_std=ss.numpy(), y=y.numpy(), z=z.numpy())
    print("transforms", mean.shape, float(mean[0]), float(std[82]))


def golden_sphere():
    conv = SphereConv2d(6, 8, 3, 1, 1)
    conv.weight.data = O.det_tensor("sphere3.weight", (8, 6, 3, 3), 31)
    conv.bias.data = O.det_tensor("sphere3.bias", (8,), 31)
    x = seeded((2, 6, 7, 10), 107)
    y3 = conv(x)
    dw = SphereConv2d(6, 6, 5, 1, 2, groups=6, bias=False)
    dw.weight.data = O.det_tensor("sphere5.weight", (6, 1, 5, 5), 31)
    y5 = dw(x)
    np.savez(os.path.join(OUT, "sphere_conv.npz"), y3=y3.numpy(), y5=y5.numpy())


def golden_embeddings():
    ts = torch.tensor([2018010100, 2020022906, 2019123118, 2016070112])
    ye = get_year_sincos_embedding(ts, embedding_dim=256)
    cfg = O.denoiser_config("375M")
    rope = LaDCastRotaryPosEmbed_from_grid(cfg["rope_axes_dim"], [1, 1, 1], theta=cfg["rope_theta"])
    lat = torch.linspace(float(np.deg2rad(-499.5)), float(np.deg2rad(508.5)), 15)
    lon = torch.linspace(float(np.deg2rad(5.25)), float(np.deg2rad(353.25)), 30)
    cos_p, sin_p = rope(torch.zeros(1, 84, 4, 15, 30), [torch.arange(1, 5).float(), lat, lon])
    cos_c, sin_c = rope(torch.zeros(1, 84, 1, 15, 30), [torch.arange(0, 1).float(), lat, lon])
    rows = np.arange(0, 1800, 37)
    np.savez(os.path.join(OUT, "embeddings.npz"), ts=ts.numpy(), year=ye.numpy(), rows=rows,
             cos_p=cos_p[rows].numpy(), sin_p=sin_p[rows].numpy(), cos_c=cos_c[::9].numpy(), sin_c=sin_c[::9].numpy(),
             cos_p_colsum=cos_p.double().sum(0).numpy(), sin_p_colsum=sin_p.double().sum(0).numpy())


def golden_metrics():
    """evaluate/evaluate_ens_gpu.py:339-415 transcribed around the reference's own pointwise functions."""
    M, C, T, H, W = 5, 84, 2, 120, 16
    dec = seeded((M, C, T, H, W), 108)
    ref = seeded((C, T, H, W), 109)
    ref[82, :, 5:9, 3:7] = float("nan")
    lat_weight = torch.from_numpy(get_normalized_lat_weights_based_on_cos(np.linspace(-88.5, 90, 120)))
    SST = 82
    tabs = {k: torch.zeros(C, T, dtype=torch.float64) for k in ("ens_mse", "crps_skill", "crps_spread", "crps")}
    for t in range(T):
        dec_t, ref_t = dec[:, :, t], ref[:, t]
        weights = lat_weight.view(1, -1, 1)
        mean_t = dec_t.mean(dim=0)
        se_t = (mean_t - ref_t) ** 2 * weights
        spread_t = pointwise_crps_spread(dec_t, ensemble_dim=0) * weights
        skill_t = pointwise_crps_skill(dec_t, ref_t.unsqueeze(0), 0) * weights
        crps_t = skill_t - 0.5 * spread_t
        for name, v in (("ens_mse", se_t), ("crps_spread", spread_t), ("crps_skill", skill_t), ("crps", crps_t)):
            tabs[name][:SST, t] = v[:SST].mean(dim=(1, 2))
            tabs[name][SST : SST + 1, t] = torch.nanmean(v[SST : SST + 1], dim=(1, 2))
            tabs[name][SST + 1 :, t] = v[SST + 1 :].mean(dim=(1, 2))
    np.savez(os.path.join(OUT, "metrics.npz"), lat_weight=lat_weight.numpy(), **{k: v.numpy() for k, v in tabs.items()})


def golden_acc():
    """The reference's own get_acc (evaluate/utils.py:122-149): lat-weighted and unweighted, NaNs in truth (SST)."""
    f, t, c = seeded((84, 120, 24), 21), seeded((84, 120, 24), 22), seeded((84, 120, 24), 23, 0.3)
    t[82, :4] = float("nan")
    w = torch.from_numpy(get_normalized_lat_weights_based_on_cos(np.linspace(-88.5, 90, 120))).view(-1, 1)
    np.savez(os.path.join(OUT, "acc.npz"), weighted=get_acc(f, t, c, w).numpy(), unweighted=get_acc(f, t, c).numpy())
    print("acc", float(get_acc(f, t, c)[0]))


def golden_heun8():
    """edm_AR_sampler (Heun, fp64 state), N = 8 -> 15 denoiser calls, tiny denoiser, ensemble of 2."""
    cfg, m = build_denoiser("tiny", 11)
    pipe = AutoRegressive2DPipeline(m, EDMDPMSolverMultistepScheduler())
    known = seeded((1, 84, 1, 15, 30), 102, 0.5)
    s = ensemble_AR_sampler(pipe, sample_size=2, return_seq_len=1, num_inference_steps=8, known_latents=known,
                            timestamps=torch.tensor([2018010100]), sampler_type="edm", device="cpu")
    np.savez(os.path.join(OUT, "heun8_tiny.npz"), edm_8=s.numpy().astype(np.float32))
    print("heun8", float(s.abs().mean()))


def churn_noise_fn(seed=777):
    """randn_like for the churn goldens: fp64 draws from one seeded CPU generator, in call order."""
    g = torch.Generator("cpu").manual_seed(seed)
    return lambda x: torch.randn(x.shape, generator=g, dtype=x.dtype).to(x.device)


CHURN_KW = dict(deterministic=False, S_churn=4.0, S_min=0.05, S_max=50.0, S_noise=1.003)


def golden_heun_churn():
    """edm_AR_sampler with stochastic churn (deterministic=False, edm_sampler.py:67-76): N = 6, tiny denoiser,
    ensemble of 2; S_min / S_max chosen so that some steps have gamma = 0 and some gamma > 0."""
    cfg, m = build_denoiser("tiny", 11)
    pipe = AutoRegressive2DPipeline(m, EDMDPMSolverMultistepScheduler())
    known = seeded((1, 84, 1, 15, 30), 102, 0.5)
    s = ensemble_AR_sampler(pipe, sample_size=2, return_seq_len=1, num_inference_steps=6, known_latents=known,
                            timestamps=torch.tensor([2018010100]), sampler_type="edm", device="cpu",
                            sampler_kwargs=dict(CHURN_KW, randn_like=churn_noise_fn()))
    np.savez(os.path.join(OUT, "heun_churn_tiny.npz"), edm_churn_6=s.numpy().astype(np.float32))
    print("heun churn", float(s.abs().mean()))


def golden_denoiser_1p6b():
    """ladcast_1.6B (d=2048, 16 heads, 5+10+3 blocks; 1,605,496,660 parameters) through the unmodified reference:
    B=1, T_out=1.  Pins the oracle port's 1.6B configuration beyond the parameter count."""
    cfg, m = build_denoiser("1.6B", 14)
    x = seeded((1, 84, 1, 15, 30), 100)
    cond = seeded((1, 84, 1, 15, 30), 101, 0.5)
    t = torch.tensor([0.2306])
    ts = torch.tensor([2018070112])
    out = m(x, t, cond, time_elapsed=ts, return_dict=False)[0]
    s, a = summary(out)
    np.savez(os.path.join(OUT, "denoiser_1p6B.npz"), salt=14, B=1, T_out=1, t=t.numpy(), ts=ts.numpy(),
             out=out.numpy().astype(np.float32), ch_sum=s, ch_abs=a)
    print("denoiser 1.6B", out.shape, float(out.abs().mean()))


ALL = {"sphere": golden_sphere, "embeddings": golden_embeddings, "metrics": golden_metrics, "dcae": golden_dcae,
       "dcae_encode": golden_dcae_encode, "transforms": golden_transforms, "samplers": golden_samplers,
       "denoiser": golden_denoiser, "acc": golden_acc, "heun8": golden_heun8, "heun_churn": golden_heun_churn, "denoiser_1p6b": golden_denoiser_1p6b}

if __name__ == "__main__":
    for name in (sys.argv[1:] or list(ALL)):
        ALL[name]()
    print("golden vectors written to", OUT)
