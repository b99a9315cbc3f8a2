"""Drop-in for `ladcast.pipelines.edm_sampler.edm_AR_sampler` (reference pipelines/edm_sampler.py:11-120): EDM Heun
sampler, sampler state in float64, 2N-1 denoiser calls.  The fp64 predictor/corrector updates run in
`lc_sched_heun_step`; the stochastic-churn branch (deterministic=False, edm_sampler.py:67-76) adds
`lc_sched_heun_churn` before each predictor (noise drawn by the caller's `randn_like` on the fp64 state)."""
from typing import List, Optional, Union

import torch

from .. import _lib
from .utils import randn_tensor


@torch.no_grad()
def edm_AR_sampler(net, noise_scheduler, batch_size=1, return_seq_len=1, randn_like=torch.randn_like,
                   num_inference_steps=18, S_churn=0, S_min=0, S_max=float("inf"), S_noise=0, deterministic=True,
                   known_latents=None, timestamps: Optional[torch.LongTensor] = None,
                   generator: Optional[Union[torch.Generator, List[torch.Generator]]] = None, device="cpu"):
    if isinstance(generator, list) and len(generator) != batch_size:
        raise ValueError(
            f"You have passed a list of generators of length {len(generator)}, but requested an effective batch"
            f" size of {batch_size}. Make sure the batch size matches the length of the generators.")
    assert known_latents is not None, "known_latents must be provided"
    if isinstance(device, str):
        device = torch.device(device)
    shape = (batch_size, net.config.out_channels, return_seq_len, *known_latents.shape[-2:])
    latents = randn_tensor(shape, generator=generator, device=device, dtype=net.dtype)
    noise_scheduler.set_timesteps(num_inference_steps, device=device)
    sig32 = noise_scheduler.sigmas.to(torch.float32).cpu()
    t_steps = sig32.to(torch.float64)  # CPU, float32 values widened
    sd = float(noise_scheduler.config.sigma_data)
    lib = _lib.load()

    def coef(t):
        t32 = t.to(torch.float32)
        c_in = 1 / ((t32**2 + sd**2) ** 0.5)
        c_skip = sd**2 / (t32**2 + sd**2)
        c_out = t32 * sd / (t32**2 + sd**2) ** 0.5
        return float(c_in), float(c_skip), float(c_out), (0.25 * torch.log(t32)).reshape(1)

    latents = latents.contiguous()
    x = torch.empty(shape, device=device, dtype=torch.float64)
    x_hat, d_cur = torch.empty_like(x), torch.empty_like(x)
    x_in = torch.empty(shape, device=device, dtype=torch.float32)
    n = num_inference_steps
    with net.cached_conditioning(known_latents, timestamps, t_out=return_seq_len):
        c_in, c_skip, c_out, c_noise = coef(t_steps[0])
        # x = float64(noise) * t_0 ; x_in = float32(x * c_in(t_0))   (edm_sampler.py:44-58)
        _lib.check(lib.lc_sched_heun_init(_lib.ptr(latents), _lib.ptr(x), _lib.ptr(x_in), x.numel(), float(t_steps[0]),
                                          c_in, _lib.stream()), "lc_sched_heun_init")
        for i in range(n):
            t_cur, t_next = float(t_steps[i]), float(t_steps[i + 1])
            if not deterministic:
                # increase the noise level temporarily: t_hat, gamma and the noise scale in float32 like the reference
                gamma = min(S_churn / num_inference_steps, 2.0**0.5 - 1) if S_min <= float(sig32[i]) <= S_max else 0
                t_hat = sig32[i] + gamma * sig32[i]
                k = float((t_hat**2 - sig32[i] ** 2).sqrt() * S_noise)
                eps = randn_like(x).contiguous()  # fp64; drawn every step (also when gamma == 0) like the reference
                c_in, c_skip, c_out, c_noise = coef(t_hat)
                _lib.check(lib.lc_sched_heun_churn(_lib.ptr(x), _lib.ptr(eps), _lib.ptr(x_in), x.numel(), k, c_in,
                                                   _lib.stream()), "lc_sched_heun_churn")
                t_cur = float(t_hat)
            f = net(x_in, c_noise.to(device), known_latents, time_elapsed=timestamps).sample
            second = i < n - 1
            # last step (t_next = 0): Euler only; x_in then receives float32(x), the sampler's return value
            c_in_n, c_skip_n, c_out_n, c_noise_n = coef(t_steps[i + 1]) if second else (1.0, 0.0, 0.0, None)
            _lib.check(lib.lc_sched_heun_step(_lib.ptr(f), _lib.ptr(x), _lib.ptr(x_hat), _lib.ptr(d_cur), _lib.ptr(x_in),
                                              x.numel(), 0, t_cur, t_next, c_skip, c_out, c_in_n, _lib.stream()),
                       "lc_sched_heun_step")
            if second:
                f2 = net(x_in, c_noise_n.to(device), known_latents, time_elapsed=timestamps).sample
                _lib.check(lib.lc_sched_heun_step(_lib.ptr(f2), _lib.ptr(x), _lib.ptr(x_hat), _lib.ptr(d_cur),
                                                  _lib.ptr(x_in), x.numel(), 1, t_cur, t_next, c_skip_n, c_out_n, c_in_n,
                                                  _lib.stream()), "lc_sched_heun_step")
                c_in, c_skip, c_out, c_noise = c_in_n, c_skip_n, c_out_n, c_noise_n
    return x_in  # == x.float() (written by the last step kernel)
