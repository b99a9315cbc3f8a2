"""Workload for the per-kernel ncu captures of round 2 (profiles/r02_*.md): every kernel class of the hot path at its
production shape, inside cudaProfilerStart/Stop windows so that warm-up launches are not captured.

  ncu --set full --clock-control none --import-source on --profile-from-start off \
      -k regex:'<names>' -o gpurun_out/r02_<tag> python tools/prof_all.py <parts>

parts (any of): den (one 375M denoiser call, B=20, T_out=4 + scheduler step + AR feedback), dec (80-frame DC-AE decode),
met (metrics kernel, ens 20 and 50, 84x4 planes), heun (fp64 Heun steps), attn16 (attention B=13, H=16: the 1.6B shape).
Also writes gpurun_out/prof_all_classes.json: algorithmic bytes / FLOPs per launch and class (lc_prof_collect_all)."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import DCAE_KW, denoiser_kwargs
from ladcast_b200 import _lib
from ladcast_b200.models import AutoencoderDC, LaDCastTransformer3DModel

parts = sys.argv[1:] or ["den", "dec", "met", "heun"]
lib = _lib.load()
cudart = torch.cuda.cudart()
dev = torch.device("cuda", 0)
torch.manual_seed(1234)
B = int(os.environ.get("PROF_B", "20"))
lib.lc_prof_enable(1)


def window(fn):
    torch.cuda.synchronize()
    cudart.cudaProfilerStart()
    fn()
    torch.cuda.synchronize()
    cudart.cudaProfilerStop()


if "den" in parts:
    m = LaDCastTransformer3DModel.from_config(denoiser_kwargs("375M")).to(dev)
    x = torch.randn(B, 84, 4, 15, 30, device=dev)
    cond = torch.randn(B, 84, 1, 15, 30, device=dev) * 0.5
    t = torch.full((1,), 0.5, device=dev)
    ts = torch.tensor([2018010100])
    x0p, xin, known, img = torch.zeros_like(x), torch.empty_like(x), torch.empty_like(cond), torch.randn_like(x)
    mean, std = torch.zeros(84, device=dev), torch.ones(84, device=dev)
    with m.cached_conditioning(cond, ts, t_out=4):
        for _ in range(2):
            f = m(x, t, cond, time_elapsed=ts).sample

        def den():
            f = m(x, t, cond, time_elapsed=ts).sample
            _lib.check(lib.lc_sched_scale_input(_lib.ptr(img), _lib.ptr(xin), img.numel(), 0.3, _lib.stream()), "scale")
            _lib.check(lib.lc_sched_dpmpp2m_step(_lib.ptr(f), _lib.ptr(img), _lib.ptr(x0p), _lib.ptr(xin), img.numel(), 0.1, 0.9,
                                                 0.5, 0.5, 0.25, 0.7, _lib.stream()), "sched")
            _lib.check(lib.lc_latent_feedback(_lib.ptr(img), _lib.ptr(known), None, _lib.ptr(mean), _lib.ptr(std), 0.5, B, 84, 4,
                                              1, 450, _lib.stream()), "feedback")

        window(den)
    m._release()

if "attn16" in parts:
    qkv = torch.randn(13, 2250, 3 * 2048, device=dev).bfloat16()
    o = torch.empty(13, 2250, 2048, device=dev, dtype=torch.bfloat16)
    for _ in range(2):
        lib.lc_attention(_lib.PRECISION_BF16, _lib.ptr(qkv), _lib.ptr(o), 13, 2250, 16, _lib.stream())
    window(lambda: lib.lc_attention(_lib.PRECISION_BF16, _lib.ptr(qkv), _lib.ptr(o), 13, 2250, 16, _lib.stream()))

if "dec" in parts:
    ae = AutoencoderDC(**DCAE_KW).to(dev)
    lat = torch.randn(B, 84, 4, 15, 30, device=dev)
    mean, std = torch.zeros(84, device=dev), torch.ones(84, device=dev)
    out = torch.empty(B, 84, 4, 120, 240, device=dev)
    ae.decode_ens_fused(lat, mean, std, latent_mean=mean, latent_std=std, out=out)
    window(lambda: ae.decode_ens_fused(lat, mean, std, latent_mean=mean, latent_std=std, out=out))
    ae._release()

if "met" in parts:
    from ladcast_b200.evaluate.utils import ensemble_metrics

    truth = torch.randn(84, 4, 120, 240, device=dev)
    for M in (20, 50):
        f = torch.randn(M, 84, 4, 120, 240, device=dev)
        ensemble_metrics(f, truth)
        window(lambda: ensemble_metrics(f, truth))
        del f

if "heun" in parts:
    n = B * 84 * 4 * 450
    f = torch.randn(n, device=dev)
    xd, xh, dc = [torch.randn(n, device=dev, dtype=torch.float64) for _ in range(3)]
    xin = torch.empty(n, device=dev)

    def heun():
        for ph in (0, 1):
            _lib.check(lib.lc_sched_heun_step(_lib.ptr(f), _lib.ptr(xd), _lib.ptr(xh), _lib.ptr(dc), _lib.ptr(xin), n, ph, 2.0,
                                              1.5, 0.1, 0.9, 0.6, _lib.stream()), "heun")

    heun()
    window(heun)

torch.cuda.synchronize()
cls = _lib.prof_collect()
per = {k: {"launches": v["launches"], "bytes_per_launch": v["bytes"] / v["launches"], "flops_per_launch": v["flops"] / v["launches"]}
       for k, v in cls.items()}
os.makedirs("gpurun_out", exist_ok=True)
json.dump(per, open("gpurun_out/prof_all_classes.json", "w"), indent=1)
print("prof_all done:", parts)
