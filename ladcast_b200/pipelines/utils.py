"""Drop-ins for the hot-path helpers of `ladcast.pipelines.utils` (reference pipelines/utils.py): `ensemble_AR_sampler`
(:665-742), `decode_latent_ens` (:52-80), `Fields2DPipelineOutput` (:26-35), plus `randn_tensor` (diffusers
utils/torch_utils.py) and the member-sharding helper the multi-GPU driver uses."""
import copy
from dataclasses import dataclass
from typing import Optional, Sequence, Union

import numpy as np
import torch


@dataclass
class Fields2DPipelineOutput:
    fields: Union[torch.Tensor, np.ndarray]

    def __getitem__(self, i):
        return (self.fields,)[i]


_NOISE_STAGING = {}  # (shape, dtype) -> [pinned host buffer, event of the last H2D copy out of it]


def randn_tensor(shape, generator=None, device=None, dtype=None):
    """Per-sample CPU generators -> one draw of (1, *shape[1:]) each (diffusers randn_tensor semantics), then moved to
    `device`.  For a CUDA target the draws are written into a reused pinned staging buffer and copied with an
    asynchronous H2D, so the host is not blocked behind the previous AR step's kernels (a pageable multi-MB copy is)."""
    device = torch.device(device) if device is not None else torch.device("cpu")
    if isinstance(generator, list) and len(generator) == 1:
        generator = generator[0]
    if isinstance(generator, list):
        one = (1,) + tuple(shape[1:])
        if device.type != "cuda":
            parts = [torch.randn(one, generator=g, device=g.device, dtype=dtype) for g in generator]
            return torch.cat(parts, dim=0).to(device)
        key = ((len(generator),) + tuple(shape[1:]), dtype)
        slot = _NOISE_STAGING.get(key)
        if slot is None:
            slot = [torch.empty(key[0], dtype=dtype, pin_memory=True), None]
            _NOISE_STAGING[key] = slot
        if slot[1] is not None:
            slot[1].synchronize()  # the previous copy out of this buffer (issued one AR step ago) has long finished
        for i, g in enumerate(generator):
            torch.randn(one, generator=g, dtype=dtype, out=slot[0][i : i + 1])
        dst = slot[0].to(device, non_blocking=True)
        slot[1] = torch.cuda.Event()
        slot[1].record(torch.cuda.current_stream(device))
        return dst
    gdev = generator.device if generator is not None else device
    return torch.randn(tuple(shape), generator=generator, device=gdev, dtype=dtype).to(device)


def member_shard(sample_size: int, rank: int, world_size: int) -> range:
    """Contiguous block of global member indices owned by `rank`: [ceil(r*M/W), ceil((r+1)*M/W))."""
    lo = -(-rank * sample_size // world_size)
    hi = -(-(rank + 1) * sample_size // world_size)
    return range(lo, hi)


@torch.no_grad()
def ensemble_AR_sampler(pipeline, sample_size: int, return_seq_len: int, num_inference_steps: int, sampler_kwargs=None,
                        known_latents: torch.Tensor = None, timestamps: Optional[torch.LongTensor] = None,
                        batch_size: int = 64, sampler_type: Optional[str] = "edm", device="cpu",
                        member_indices: Optional[Sequence[int]] = None):
    """timestamps: (1,) or (B,) int YYYYMMDDHH; known_latents: (1 or B, C, T, H, W); returns (sample_size, C, T, H, W).
    Member m always draws its noise from torch.Generator('cpu').manual_seed(m) — at every AR step (reference quirk).
    `member_indices` (extension) selects which GLOBAL members this process samples (default 0..sample_size-1), so
    that ranks of a multi-GPU job reproduce exactly the members a single process would."""
    from .edm_sampler import edm_AR_sampler

    if member_indices is None:
        member_indices = list(range(sample_size))
    member_indices = list(member_indices)
    assert len(member_indices) == sample_size, "member_indices must list sample_size members"
    sizes = [batch_size] * (sample_size // batch_size)
    if sample_size % batch_size:
        sizes.append(sample_size % batch_size)
    samples = None
    sampler_kwargs = sampler_kwargs or {}
    scheduler = copy.deepcopy(pipeline.scheduler) if sampler_type == "edm" else pipeline.scheduler
    count = 0
    for n in sizes:
        gens = [torch.Generator("cpu").manual_seed(int(m) % (1 << 32)) for m in member_indices[count : count + n]]
        if known_latents.shape[0] == 1:
            known = known_latents.expand(n, *known_latents.shape[1:]).contiguous()
        else:
            known = known_latents[count : count + n] if known_latents.shape[0] == sample_size else known_latents
        if sampler_type == "edm":
            out = edm_AR_sampler(pipeline.ar_model, scheduler, batch_size=n, return_seq_len=return_seq_len,
                                 num_inference_steps=num_inference_steps, generator=gens, device=device,
                                 known_latents=known, timestamps=timestamps, **sampler_kwargs)
        elif sampler_type == "pipeline":
            out = pipeline(batch_size=n, return_seq_len=return_seq_len, num_inference_steps=num_inference_steps,
                           generator=gens, known_latents=known, timestamps=timestamps, return_dict=False,
                           do_edm_style=True, **sampler_kwargs)[0]
        else:
            raise ValueError(f"unknown sampler_type {sampler_type!r}")
        if n == sample_size:
            return out  # one batch: the sampler's own output buffer, no gather copy
        if samples is None:
            samples = torch.empty(sample_size, *out.shape[1:], device=out.device, dtype=out.dtype)
        samples[count : count + n] = out
        count += n
    return samples


@torch.no_grad()
def decode_latent_ens(encdec_model, latents: torch.Tensor, mean_tensor: Optional[torch.Tensor] = None,
                      std_tensor: Optional[torch.Tensor] = None, extract_first: Optional[int] = None) -> torch.Tensor:
    """latents: (B, C, T, H, W) -> decoded fields (B, 84, T, 8H, 8W), de-normalised with mean/std if given."""
    B, C, T, H, W = latents.shape
    if extract_first is None:
        extract_first = T
    if hasattr(encdec_model, "decode_ens_fused"):
        # native path: 5-D in, 5-D out, de-normalisation in the last epilogue — no permute / reshape copies
        return encdec_model.decode_ens_fused(latents, mean_tensor, std_tensor, extract_first=extract_first)
    z = latents[:, :, :extract_first].to(encdec_model.device).permute(0, 2, 1, 3, 4).reshape(B * extract_first, C, H, W)
    y = encdec_model.decode(z.contiguous()).sample
    if mean_tensor is not None:
        y = y * std_tensor.to(y.device)[None, :, None, None] + mean_tensor.to(y.device)[None, :, None, None]
    y = y.reshape(B, extract_first, *y.shape[1:]).permute(0, 2, 1, 3, 4)
    return y


def advance_timestamp(stamp: int, hours: int) -> int:
    """YYYYMMDDHH + hours -> YYYYMMDDHH (roll_out_serial's `current_time + Timedelta(...)`, pipelines/utils.py:538-541)."""
    from datetime import datetime, timedelta

    s = str(int(stamp))
    t = datetime(int(s[:4]), int(s[4:6]), int(s[6:8]), int(s[8:10])) + timedelta(hours=hours)
    return int(t.strftime("%Y%m%d%H"))


@torch.no_grad()
def rollout_step(pipeline, encdec_model, known: torch.Tensor, stamp: torch.Tensor, ensemble_size: int,
                 latent_mean: torch.Tensor, latent_std: torch.Tensor, field_mean: Optional[torch.Tensor],
                 field_std: Optional[torch.Tensor], num_inference_steps: int = 20, return_seq_len: int = 4,
                 sampler_type: str = "pipeline", member_indices: Optional[Sequence[int]] = None,
                 return_latent: bool = False, target_std: float = 0.5, t_in: Optional[int] = None,
                 out: Optional[torch.Tensor] = None, extract_first: Optional[int] = None):
    """One AR step of `roll_out_serial` (reference pipelines/utils.py:533-585), entirely in library kernels: sample
    T_out lead steps for every member, feed the last T_in frames back (lc_latent_feedback), then either de-normalise
    the latents (return_latent) or decode them — the latent de-normalisation runs inside the decoder's first kernel and
    the field de-normalisation in its last epilogue, which writes (ens, C, T_out, H, W) directly.
    All statistics tensors must already live on the device.  Returns (result, next known latents)."""
    from .. import _lib

    dev = pipeline._execution_device
    samples = ensemble_AR_sampler(pipeline, sample_size=ensemble_size, return_seq_len=return_seq_len,
                                  num_inference_steps=num_inference_steps, known_latents=known, timestamps=stamp,
                                  sampler_type=sampler_type, device=dev, member_indices=member_indices).contiguous()
    B, C, T, h, w = samples.shape
    t_in = known.shape[2] if t_in is None else t_in
    known_next = torch.empty((B, C, t_in, h, w), device=dev, dtype=torch.float32)
    phys = torch.empty_like(samples) if return_latent else None
    with torch.cuda.device(dev):
        _lib.check(_lib.load().lc_latent_feedback(_lib.ptr(samples), _lib.ptr(known_next), _lib.ptr(phys),
                                                  _lib.ptr(latent_mean), _lib.ptr(latent_std), float(target_std), B, C, T,
                                                  t_in, h * w, _lib.stream()), "lc_latent_feedback")
    if return_latent:
        return (phys if extract_first in (None, T) else phys[:, :, :extract_first]), known_next
    fields = encdec_model.decode_ens_fused(samples, field_mean, field_std, extract_first=extract_first,
                                           latent_mean=latent_mean, latent_std=latent_std, target_std=target_std, out=out)
    return fields, known_next


@torch.no_grad()
def roll_out_latent(pipeline, encdec_model, known_latents: torch.Tensor, init_timestamp: int, ensemble_size: int,
                    latent_mean: torch.Tensor, latent_std: torch.Tensor, field_mean: Optional[torch.Tensor] = None,
                    field_std: Optional[torch.Tensor] = None, num_inference_steps: int = 20, return_seq_len: int = 4,
                    total_lead_time_hour: int = 240, step_size_hour: int = 6, sampler_type: str = "pipeline",
                    member_indices: Optional[Sequence[int]] = None, return_latent: bool = False, target_std: float = 0.5,
                    out: Optional[torch.Tensor] = None, max_ar_steps: Optional[int] = None,
                    return_ensemble_mean: bool = False):
    """Tensor-in / tensor-out core of `roll_out_serial` (reference pipelines/utils.py:533-654): the AR loop
    [sample -> feed the last T_in frames back -> de-normalise -> decode] from already encoded, normalised
    `known_latents` (1, C, T_in, h, w) given on the HOST or the device.

    Returns a pinned HOST tensor of shape (n_ar_steps, ensemble, C, T_out, H, W): block s holds lead steps
    s*T_out+1 .. (s+1)*T_out (decoded fields in physical units, or de-normalised latents when return_latent).
    `rollout_as_lead_major(out)` gives the reference's (ensemble, C, lead, H, W) view.  Each AR step's block is one
    contiguous asynchronous device->host copy on a side stream that overlaps the next AR step's compute.
    `member_indices`: global member ids owned by this process (multi-GPU member sharding).
    `return_ensemble_mean` (reference :299, :608-630): only the ensemble mean of every decoded block leaves the device
    (shape (n_ar_steps, 1, C, T_out, H, W)); the mean over members is taken by the metrics kernel's pointwise mode.
    When T_out does not divide the number of lead steps, the last block's surplus frames are not decoded (the
    reference's `pred_selection`, :536-537) and stay NaN in the output."""
    import math

    if return_latent and return_ensemble_mean:
        raise ValueError("return_ensemble_mean must be False when return_latent is True.")
    dev = pipeline._execution_device
    total = total_lead_time_hour // step_size_hour
    if total_lead_time_hour % step_size_hour != 0:
        raise ValueError("total_lead_time_hour must be divisible by step_size_hour.")
    reps = math.ceil(total / return_seq_len)
    if max_ar_steps is not None:
        reps = min(reps, max_ar_steps)
    t_in = known_latents.shape[2]
    known = known_latents.to(dev, torch.float32, non_blocking=True)
    lm = latent_mean.to(dev, torch.float32).contiguous()
    ls = latent_std.to(dev, torch.float32).contiguous()
    fm = field_mean.to(dev, torch.float32).contiguous() if field_mean is not None else None
    fs = field_std.to(dev, torch.float32).contiguous() if field_std is not None else None
    copy_stream = torch.cuda.Stream(device=dev)
    for step in range(reps):
        stamp = torch.tensor([advance_timestamp(init_timestamp, step * step_size_hour * return_seq_len)])
        sel = min(return_seq_len, total - step * return_seq_len)  # lead steps of this block that exist (pred_selection)
        res, known = rollout_step(pipeline, encdec_model, known, stamp, ensemble_size, lm, ls, fm, fs,
                                  num_inference_steps=num_inference_steps, return_seq_len=return_seq_len,
                                  sampler_type=sampler_type, member_indices=member_indices, return_latent=return_latent,
                                  target_std=target_std, t_in=t_in, extract_first=sel)
        if return_ensemble_mean:
            from ..evaluate.utils import ensemble_mean

            res = ensemble_mean(res).unsqueeze(0)  # (1, C, sel, H, W), reduced on the device
        if out is None:
            shape = (reps, res.shape[0], res.shape[1], return_seq_len, *res.shape[3:])
            out = torch.empty(shape, dtype=torch.float32, pin_memory=True)
            if total % return_seq_len:
                out[-1].fill_(float("nan"))
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream(dev))
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done)
            if sel == return_seq_len:
                out[step].copy_(res, non_blocking=True)
            else:
                out[step][:, :, :sel].copy_(res, non_blocking=True)
        res.record_stream(copy_stream)  # the allocator may reuse the block only after the copy has drained
    copy_stream.synchronize()
    return out


def encode_initial_condition(encdec_model, input_fields: torch.Tensor, static_fields: Optional[torch.Tensor],
                             latent_mean: torch.Tensor, latent_std: torch.Tensor, target_std: float = 0.5) -> torch.Tensor:
    """The encode + normalise front of `roll_out_serial` (reference pipelines/utils.py:457-481 and
    normalize_transform_3D, dataloader/utils.py:223-231): standardised fields (C, T_in, H, W) + static channels
    (C_s, H, W) -> normalised known latents (1, C_lat, T_in, h, w) on the autoencoder's device."""
    if input_fields.dim() != 4:
        raise ValueError("input_fields must be (C, T_in, H, W)")
    x = input_fields.permute(1, 0, 2, 3)  # (T_in, C, H, W): frames are the encoder's batch
    st = None
    if static_fields is not None:
        st = static_fields.unsqueeze(0).expand(x.shape[0], -1, -1, -1)
    z = encdec_model.encode_fused(x, latent_mean, latent_std, target_std, static_conditioning_tensor=st)
    return z.permute(1, 0, 2, 3).unsqueeze(0).contiguous()


def roll_out_serial(pipeline, encdec_model, input_fields: torch.Tensor, static_fields: Optional[torch.Tensor],
                    init_timestamp, ensemble_size: int, latent_mean: torch.Tensor, latent_std: torch.Tensor,
                    field_mean: Optional[torch.Tensor] = None, field_std: Optional[torch.Tensor] = None,
                    noise_level: float = 0.0, generator: Optional[torch.Generator] = None,
                    reference_layout: bool = False, raw_fields: Optional[torch.Tensor] = None, **kw):
    """Tensor-in / tensor-out `roll_out_serial` (reference pipelines/utils.py:250-661 without the xarray layer): encode
    the initial condition (+ static channels), normalise it, optionally perturb it, then run `roll_out_latent`.

    input_fields: standardised fields of the T_in input times, (C, T_in, H, W) for one forecast or
    (n_init, C, T_in, H, W) with `init_timestamp` a sequence for the reference's loop over init times (:445).
    noise_level (:518-528): known latents += randn * noise_level * latent_std (drawn from `generator` on the host).
    Keyword options of `roll_out_latent` pass through (num_inference_steps, return_seq_len, total_lead_time_hour,
    return_latent, return_ensemble_mean, sampler_type, member_indices, ...).

    Returns (rollout, known_latents) — the AR-blocked host tensor of `roll_out_latent` (stacked over init times when
    several are given) — or, with reference_layout=True, the reference's `return_tensor=True` result: a tensor
    (n_init, return_size, C, lead+1, H, W) whose slot 0 holds the un-normalised input field of the forecast time
    (`raw_fields`, (n_init,) C, H, W; NaN if not given) or, for return_latent, the encoded initial latent (:466-497)."""
    multi = input_fields.dim() == 5
    fields = input_fields if multi else input_fields.unsqueeze(0)
    stamps = list(init_timestamp) if multi else [init_timestamp]
    if len(stamps) != fields.shape[0]:
        raise ValueError("one init_timestamp per forecast is required")
    target_std = kw.get("target_std", 0.5)
    outs, knowns = [], []
    for i, stamp in enumerate(stamps):
        known = encode_initial_condition(encdec_model, fields[i], static_fields, latent_mean, latent_std, target_std)
        knowns.append(known)
        start = known
        if noise_level and noise_level > 0:
            noise = torch.randn(known.shape, generator=generator, dtype=torch.float32)
            start = known + (noise * noise_level * latent_std.to(torch.float32).reshape(1, -1, 1, 1, 1)).to(known.device)
        outs.append(roll_out_latent(pipeline, encdec_model, start, int(stamp), ensemble_size, latent_mean, latent_std,
                                    field_mean, field_std, **kw))
    if not reference_layout:
        if multi:
            return torch.stack(outs, dim=0), torch.cat(knowns, dim=0)
        return outs[0], knowns[0]
    total = kw.get("total_lead_time_hour", 240) // kw.get("step_size_hour", 6)
    res = []
    for i, o in enumerate(outs):
        lead = rollout_as_lead_major(o, total)  # (return_size, C, total, H, W)
        first = torch.full_like(lead[:, :, :1], float("nan"))
        if kw.get("return_latent", False):
            z0 = (knowns[i][0, :, -1].cpu() / target_std) * latent_std.reshape(-1, 1, 1) + latent_mean.reshape(-1, 1, 1)
            first[:] = z0[None, :, None]
        elif raw_fields is not None:
            rf = raw_fields[i] if raw_fields.dim() == 4 else raw_fields
            first[:] = rf.to("cpu", torch.float32)[None, :, None]
        res.append(torch.cat([first, lead], dim=2))
    return torch.stack(res, dim=0)


def save_latents_npy(path_dir: str, init_timestamp: int, initial_latent: torch.Tensor, rollout: torch.Tensor) -> str:
    """Writes `latent_YYYYMMDDHH.npy` in the reference's on-disk format (evaluate/pred_rollout.py:421-430): float32
    array (ensemble, C, T+1, h, w) whose t=0 slot holds the encoded initial condition (de-normalised latent
    (C, h, w), same for every member) and t>=1 the predicted latents.  `rollout` is roll_out_latent(...,
    return_latent=True) output, either lead-major (ens, C, T, h, w) or AR-blocked (n_ar, ens, C, T_out, h, w)."""
    import os

    lat = rollout_as_lead_major(rollout) if rollout.dim() == 6 else rollout
    ens, C, T, h, w = lat.shape
    arr = np.empty((ens, C, T + 1, h, w), dtype=np.float32)
    arr[:, :, 0] = initial_latent.detach().to("cpu", torch.float32).numpy()[None]
    arr[:, :, 1:] = lat.detach().to("cpu", torch.float32).numpy()
    os.makedirs(path_dir, exist_ok=True)
    out = os.path.join(path_dir, f"latent_{int(init_timestamp):010d}.npy")
    np.save(out, arr)
    return out


def rollout_as_lead_major(out: torch.Tensor, n_lead: Optional[int] = None) -> torch.Tensor:
    """(n_ar, ens, C, T_out, H, W) -> the reference's (ens, C, n_ar*T_out, H, W) ordering (a copy), optionally cut to
    the first n_lead lead steps (roll_out_serial's `pred_selection` when T_out does not divide the lead count)."""
    n_ar, ens, C, T, H, W = out.shape
    y = out.permute(1, 2, 0, 3, 4, 5).reshape(ens, C, n_ar * T, H, W)
    return y if n_lead is None else y[:, :, :n_lead]
