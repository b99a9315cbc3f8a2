#!/usr/bin/env python
"""Benchmark of the LaDCast ensemble-rollout hot path (BASELINE.json metric: ensemble-member 6 h-steps / second).

  python bench.py --gpus N --steps K --warmup W            # ours: sm_100a CUDA path, one process per GPU
  python bench.py --impl reference --gpus N ...             # the reference algorithm (CPU oracle port) on host cores

One "step" = one autoregressive step of the rollout for the members resident on a GPU: `num_inference_steps`
denoiser calls of the DPM-Solver++ sampler (T_out lead steps at once), the scheduler updates, feeding the last frame
back, latent de-normalisation and the DC-AE decode of every (member, lead) frame.  It yields ens * T_out
member-6h-steps.  Prints ONE JSON line on rank 0 with

  value / ms_per_step  headline (BASELINE config 2): ladcast_375M, `--ens` members PER GPU (weak scaling: members are
                       independent, no communication inside the rollout), inputs resident in HBM;
  e2e                  the same through `roll_out_latent` with host buffers (H2D noise, D2H fields every AR step);
  roofline             tcgen05 GEMM class, live CUDA-event timing per launch; `roofline.secondary`: every other
                       kernel class (attention / sphere conv vs the bf16 peak, HBM-bound kernels vs the copy peak);
  metrics              the ensemble-metrics kernel (lat-weighted RMSE / CRPS) on the decoded fields: GB/s vs HBM peak;
  strong               BASELINE configs 4/5: ladcast_1.6B, a FIXED ensemble (20 and 50 members) sharded over the N
                       ranks with `member_shard`, plus the one collective of the path — the member->plane re-shard of
                       the decoded fields (grouped NCCL send/recv) feeding the metrics kernel — timed and checked
                       against the single-GPU metrics of the gathered fields;
  cpu_baseline         the oracle port on the host cores (bounded sample + BASELINE config 1 end to end), N=1 only.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "ensemble member 6h-steps/sec (20 denoise steps, 240h rollout)"
UNIT = "member-steps/s"

MODEL_CFG = {
    "375M": dict(num_attention_heads=12, num_layers=2, num_single_layers=4, num_refiner_layers=1),
    "1.6B": dict(num_attention_heads=16, num_layers=5, num_single_layers=10, num_refiner_layers=3),
}

# kernel classes of lc_prof_collect_all and the roofline that bounds each
TENSOR_CLASSES = ("gemm_tc", "attention_tc", "sphere_conv_tc")
CLASS_KERNELS = {
    "gemm_tc": "gemm_tc2_kernel / gemm_tc_kernel (tcgen05 bf16 GEMM, CTA-pair; denoiser linears + decoder 1x1)",
    "attention_tc": "attention_tc_persistent_kernel (tcgen05 flash attention, D=128, persistent, split P hand-over)",
    "sphere_conv_tc": "gemm_tc2_kernel in implicit-GEMM 3x3 sphere-conv mode (DC-AE)",
    "layernorm": "layernorm_kernel (LayerNorm + AdaLN modulation, fp32 in / bf16 out)",
    "qk_norm_rope": "qk_norm_rope_bf16_kernel (per-head RMSNorm(q,k) + RoPE, in place)",
    "scheduler": "dpmpp2m_kernel / scale_kernel / latent_feedback_kernel (fused scheduler step, AR feedback)",
    "dec_rmsnorm": "rmsnorm_rows_kernel (DC-AE channel RMSNorm + residual)",
    "dec_multiscale": "multiscale_fused_mma_kernel (DC-AE 5x5 depthwise on the FMA pipe + grouped 1x1 on mma.sync)",
    "dec_linear_attn": "linear_attn_mma_kernel (DC-AE ReLU linear attention on mma.sync)",
    "dec_dwconv_glu": "dwconv3_glu_kernel (DC-AE depthwise 3x3 + GLU)",
    "dec_pixel_shuffle": "pixel_shuffle_kernel (+ shortcut)",
    "dec_pad": "pad_from_* / halo_fill / in_shortcut kernels (sphere padding)",
    "metrics": "metrics_sorted_kernel (ensemble mean / CRPS skill / sorted CRPS spread, lat-weighted fp64 reduction)",
    "misc": "patchify / timestep / pooling / cast kernels",
}


def denoiser_kwargs(name):
    cfg = dict(in_channels=84, out_channels=84, attention_head_dim=128, mlp_ratio=4, patch_size=1, patch_size_t=1,
               qk_norm="rms_norm", rope_theta=256.0, rope_axes_dim=[16, 56, 56],
               rope_spatial_grid_start_pos=[-499.5, 5.25], rope_spatial_grid_end_pos=[508.5, 353.25],
               spatial_deg2rad=True, conditioning_tensor_in_channels=84,
               conditioning_tensor_rope_axes_dim=[16, 56, 56], incl_time_elapsed=True)
    cfg.update(MODEL_CFG[name])
    return cfg


DCAE_KW = dict(in_channels=89, out_channels=89, latent_channels=84, attention_head_dim=32,
               decoder_block_types=["ResBlock", "ResBlock", "EfficientViTBlock", "EfficientViTBlock"],
               decoder_block_out_channels=[252, 504, 504, 1008], decoder_layers_per_block=[4, 4, 4, 4],
               decoder_qkv_multiscales=[[], [], [5], [5]], static_channels=5)


def flops_per_member_step(name, t_out, n_denoise):
    """Algorithmic FLOPs per member-6h-step: denoiser calls + one decoded frame."""
    c = MODEL_CFG[name]
    d = c["num_attention_heads"] * 128
    n, nc, npred = 450 * (t_out + 1), 450, 450 * t_out
    blocks = c["num_layers"] + c["num_single_layers"]
    call = blocks * (24 * d * d * n + 4 * n * n * d) + c["num_refiner_layers"] * (22 * d * d * nc + 4 * nc * nc * d)
    call += 2 * 84 * d * n + 2 * d * d * nc + 2 * d * 84 * npred
    return (n_denoise * call) / t_out + 0.7814e12


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                                          "200", "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                         text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU baseline
def cpu_reference_sample(model_name, t_out, n_denoise, repeats=1, members=2):
    """The reference algorithm (oracle/ladcast_oracle.py, a CPU fp32 port validated against the unmodified reference)
    on the host cores with all threads: a bounded, BATCHED sample of the same workload — one denoiser call for
    `members` members at once (T_out lead steps, 2250 tokens) and the decode of `members` frames in one batch — scaled
    to those members' AR step (n_denoise calls + members * T_out frames)."""
    from oracle import ladcast_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.denoiser_config(model_name)
    sd = O.make_state_dict(O.denoiser_param_shapes(cfg), 1)
    acfg = O.dcae_config()
    asd = O.make_state_dict(O.dcae_decoder_param_shapes(acfg), 2)
    g = torch.Generator("cpu").manual_seed(0)
    B = members
    x = torch.randn((B, 84, t_out, 15, 30), generator=g)
    cond = torch.randn((B, 84, 1, 15, 30), generator=g) * 0.5
    z = torch.randn((B, 84, 15, 30), generator=g)
    ts = torch.tensor([2018010100])
    vals = []
    with torch.no_grad():
        O.denoiser_forward(sd, cfg, x[:1], torch.tensor([0.5]), cond[:1], ts)  # warm-up (thread pool, allocator)
        for _ in range(repeats):
            t0 = time.perf_counter()
            O.denoiser_forward(sd, cfg, x, torch.tensor([0.7]).expand(B), cond, ts)
            t_call = time.perf_counter() - t0
            t0 = time.perf_counter()
            O.dcae_decode(asd, acfg, z)
            t_dec = time.perf_counter() - t0
            vals.append((B * t_out / (n_denoise * t_call + t_out * t_dec), t_call, t_dec))
    v = sum(a for a, _, _ in vals) / len(vals)
    return {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"oracle port (CPU fp32, torch {torch.__version__}, {cores} threads): 1 batched denoiser call B={B} "
                      f"T_out={t_out} ({vals[-1][1]:.2f} s) + 1 batched decode of {B} frames ({vals[-1][2]:.2f} s), scaled to "
                      f"{n_denoise} calls + {B * t_out} frames per AR step of {B} members"}, vals


def cpu_config1():
    """BASELINE config 1 end to end through the oracle port: ladcast_375M, ensemble_size=2, num_inference_steps=5,
    1 lead step (T_out=1), synthetic 84x120x240 input through DCAE encode -> denoise -> decode, fp32, batched."""
    from oracle import ladcast_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.denoiser_config("375M")
    sd = O.make_state_dict(O.denoiser_param_shapes(cfg), 1)
    acfg = O.dcae_config()
    asd = O.make_state_dict(O.dcae_decoder_param_shapes(acfg), 2)
    asd.update(O.make_state_dict(O.dcae_encoder_param_shapes(acfg), 3))
    g = torch.Generator("cpu").manual_seed(0)
    fields = torch.randn((1, 84, 120, 240), generator=g)
    static = torch.randn((1, 5, 120, 240), generator=g)
    lat_mean, lat_std = torch.zeros(84), torch.ones(84)
    with torch.no_grad():
        t0 = time.perf_counter()
        z = O.dcae_encode(asd, acfg, fields, static)
        known = O.normalize_latent(z.permute(1, 0, 2, 3).unsqueeze(0), lat_mean, lat_std, 0.5)
        t_enc = time.perf_counter() - t0
        O.rollout(sd, cfg, asd, acfg, known, [0, 1], 2018010100, 1, 1, 5, lat_mean, lat_std, sampler="pipeline")
        dt = time.perf_counter() - t0
    return {"seconds": round(dt, 2), "encode_seconds": round(t_enc, 2), "member_steps_per_s": round(2.0 / dt, 4),
            "config": "BASELINE config 1: 375M, ens=2, 5 denoise steps, T_out=1, encode+denoise+decode, fp32, "
                      f"{cores} threads (5 denoise steps, so not comparable with the 20-step headline metric)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    base, vals = cpu_reference_sample(args.model, args.t_out, args.denoise_steps, repeats=max(1, args.steps))
    per_step_ms = 1e3 * (time.perf_counter() - t0) / max(1, args.steps)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, 1), "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(args, world):
    return {"workload": f"ladcast_{args.model} 240h rollout: AR step = {args.denoise_steps} DPM-Solver++ denoiser calls "
                        f"(T_in=1, T_out={args.t_out}, 2250 tokens) + DC-AE decode of ens*T_out frames to 84x120x240",
            "ensemble_per_gpu": args.ens, "ensemble_total": args.ens * world, "denoise_steps": args.denoise_steps,
            "t_out": args.t_out, "sampler": "pipeline (DPM-Solver++ 2M)", "parallelism": f"member-sharded x{world}",
            "l2": "working set per step (>1.5 GB of activations) exceeds the 126 MB L2; no explicit flush"}


# ------------------------------------------------------------------------------------------------ GPU legs
class Rollout:
    """Device-resident AR loop for a list of global member ids: exactly `rollout_step` of the product path."""

    def __init__(self, model_name, members, args, dev, ae=None, model=None, out=None):
        from ladcast_b200.models import AutoencoderDC, LaDCastTransformer3DModel
        from ladcast_b200.pipelines import AutoRegressive2DPipeline, EDMDPMSolverMultistepScheduler

        torch.manual_seed(1234)  # identical random-init weights on every rank
        self.model = model if model is not None else LaDCastTransformer3DModel.from_config(denoiser_kwargs(model_name)).to(dev)
        self.ae = ae if ae is not None else AutoencoderDC(**DCAE_KW).to(dev)
        self.pipe = AutoRegressive2DPipeline(self.model, EDMDPMSolverMultistepScheduler())
        self.members, self.args, self.dev = list(members), args, dev
        g = torch.Generator("cpu").manual_seed(7)
        self.known0 = torch.randn((1, 84, 1, 15, 30), generator=g) * 0.5
        self.lat_mean, self.lat_std = torch.randn(84, generator=g) * 0.1, torch.rand(84, generator=g) + 0.5
        self.fld_mean, self.fld_std = torch.randn(84, generator=g), torch.rand(84, generator=g) + 0.5
        self.stats = [t.to(dev).contiguous() for t in (self.lat_mean, self.lat_std, self.fld_mean, self.fld_std)]
        self.known = self.known0.to(dev)
        self.out = out if out is not None else (
            torch.empty((len(self.members), 84, args.t_out, 120, 240), device=dev) if self.members else None)
        self.stamp = torch.tensor([2018010100])  # the date embedding is recomputed every AR step like the reference

    def step(self):
        from ladcast_b200.pipelines.utils import rollout_step

        if not self.members:
            return None
        a = self.args
        fields, self.known = rollout_step(self.pipe, self.ae, self.known, self.stamp, len(self.members), *self.stats,
                                          num_inference_steps=a.denoise_steps, return_seq_len=a.t_out,
                                          sampler_type="pipeline", member_indices=self.members, out=self.out)
        return fields

    def probe_host(self, lib):
        """Host cost of ENQUEUEING one denoiser call into an empty stream (the launch queue of a whole AR step exceeds
        the driver's queue depth, so timing the enqueue of a full step only measures back-pressure) vs its device time."""
        B = len(self.members)
        if B == 0:
            return None
        a = self.args
        known = self.known if self.known.shape[0] == B else self.known.expand(B, -1, -1, -1, -1).contiguous()
        x = torch.randn((B, 84, a.t_out, 15, 30), device=self.dev)
        t = torch.full((1,), 0.5, device=self.dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with self.model.cached_conditioning(known, self.stamp, t_out=a.t_out):
            self.model(x, t, known, time_elapsed=self.stamp)
            torch.cuda.synchronize()
            n0 = lib.lc_launch_count()
            h0 = time.perf_counter()
            e0.record()
            self.model(x, t, known, time_elapsed=self.stamp)
            e1.record()
            host_us = 1e6 * (time.perf_counter() - h0)
            torch.cuda.synchronize()
        return {"host_enqueue_us_per_denoiser_call": round(host_us, 1),
                "device_us_per_denoiser_call": round(1e3 * e0.elapsed_time(e1), 1),
                "launches_per_denoiser_call": int(lib.lc_launch_count() - n0)}

    def release(self):
        self.model._release()
        self.model = self.pipe = None


def timed_loop(run, steps, warmup, barrier, world, dev, lib, clock_index=None):
    """W warm-up + K timed AR steps; CUDA events on the launching stream, max over ranks; also the host time spent
    ENQUEUEING the K steps (launch-bound when it approaches the device time)."""
    import torch.distributed as dist

    out = None
    for _ in range(warmup):
        out = run.step()
    barrier()
    clocks = ClockSampler(clock_index) if clock_index is not None else None
    launches0 = lib.lc_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    h0 = time.perf_counter()
    for _ in range(steps):
        out = run.step()
    host_ms = 1e3 * (time.perf_counter() - h0)
    e1.record()
    barrier()
    launches = lib.lc_launch_count() - launches0
    clk = clocks.stop() if clocks else None
    t = torch.tensor([e0.elapsed_time(e1), host_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, host_ms = [float(v) for v in t.tolist()]
    return out, ms, host_ms, int(launches), clk


def profiled_step(run, lib, _lib):
    """One more identical step with every launch bracketed by CUDA events -> per-class ms / FLOPs / bytes."""
    lib.lc_prof_enable(1)
    run.step()
    torch.cuda.synchronize()
    classes = _lib.prof_collect()
    lib.lc_prof_enable(0)
    return classes


def class_rooflines(classes, peaks, step_ms):
    """roofline objects per kernel class: tensor classes vs the sustained bf16 peak, the rest vs the HBM copy peak."""
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_gb = float(peaks.get("hbm_gbs", 6650.0))
    src = "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)"
    out = {}
    for name, c in classes.items():
        ms = c["ms"]
        if ms <= 0:
            continue
        r = {"kernel": CLASS_KERNELS.get(name, name), "launches": c["launches"], "ms": round(ms, 3),
             "avg_launch_ms": round(ms / c["launches"], 4), "share_of_step": round(ms / step_ms, 4)}
        if name in TENSOR_CLASSES:
            ach = c["flops"] / (ms * 1e-3) / 1e12
            r.update({"bound": "tensor", "achieved": round(ach, 1), "peak": peak_tf, "unit": "TFLOP/s",
                      "frac": round(ach / peak_tf, 4), "peak_source": f"{src} bf16_tflops_sustained (kernel timed inside a long step)"})
        else:
            ach = c["bytes"] / (ms * 1e-3) / 1e9
            r.update({"bound": "hbm", "achieved": round(ach, 1), "peak": peak_gb, "unit": "GB/s",
                      "frac": round(ach / peak_gb, 4), "peak_source": f"{src} hbm_gbs", "algorithmic_bytes": c["bytes"]})
        out[name] = r
    return out


def metrics_leg(fields, lib, _lib, peaks, extra_members=(50,)):
    """Ensemble-metrics kernel on the decoded fields of the last AR step ([ens, 84, T, 120, 240]) and, for the larger
    ensemble of config 4, on a synthetic block of the same shape: CUDA-event time and GB/s vs the HBM copy peak.
    Inputs (fields + truth) exceed the 126 MB L2."""
    from ladcast_b200.evaluate.utils import ensemble_metrics

    peak_gb = float(peaks.get("hbm_gbs", 6650.0))
    res = []
    g = torch.Generator("cpu").manual_seed(11)
    truth = torch.randn((84, fields.shape[2], 120, 240), generator=g).to(fields.device)
    cases = [(fields.shape[0], fields)]
    for m in extra_members:
        try:
            cases.append((m, torch.randn((m,) + tuple(fields.shape[1:]), device=fields.device)))
        except Exception:
            pass
    for m, f in cases:
        for _ in range(3):
            ensemble_metrics(f, truth)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            tabs = ensemble_metrics(f, truth)
        e1.record()
        torch.cuda.synchronize()
        call_ms = e0.elapsed_time(e1) / reps
        lib.lc_prof_enable(1)  # the reduction kernel alone (CUDA events around its launch)
        for _ in range(reps):
            ensemble_metrics(f, truth)
        torch.cuda.synchronize()
        k = _lib.prof_collect().get("metrics", {"ms": 0.0, "launches": 1})
        lib.lc_prof_enable(0)
        ms = k["ms"] / max(1, k["launches"])
        by = 4.0 * (m + 1) * truth.numel()
        res.append({"kernel": CLASS_KERNELS["metrics"], "members": m, "planes": 84 * int(fields.shape[2]),
                    "kernel_ms": round(ms, 4), "call_ms": round(call_ms, 4), "algorithmic_bytes": by,
                    "achieved": round(by / (ms * 1e-3) / 1e9, 1) if ms > 0 else None, "peak": peak_gb, "unit": "GB/s",
                    "frac": round(by / (ms * 1e-3) / 1e9 / peak_gb, 4) if ms > 0 else None, "bound": "hbm",
                    "note": "call_ms = ensemble_metrics() end to end (memsets, kernel, [C,T] table assembly)",
                    "finite": bool(all(torch.isfinite(v).all() for v in tabs.values()))})
        del f
    return res


def strong_leg(args, ens_total, rank, world, dev, ae, model, barrier, lib, _lib, steps, warmup):
    """BASELINE configs 4/5: ladcast_1.6B, `ens_total` members sharded over the ranks (member_shard), AR steps timed as
    the headline; then the one exchange of the path + metrics, timed and checked against the single-GPU result."""
    import torch.distributed as dist

    from ladcast_b200.evaluate.utils import ensemble_metrics, ensemble_metrics_distributed
    from ladcast_b200.pipelines.utils import member_shard

    members = list(member_shard(ens_total, rank, world))
    per_rank = [len(member_shard(ens_total, r, world)) for r in range(world)]
    out_buf = None
    if world > 1:  # (collective: every rank takes part, also one without members)
        # decoded fields are produced straight into this rank's peer-visible buffer (CUDA IPC), so the fused
        # peer-memory metrics below read them in place: no staging copy on either side
        try:
            from ladcast_b200.evaluate.utils import _peer_buffer

            pb = _peer_buffer(max(per_rank) * 84 * args.t_out * 120 * 240, dev)
            if members:
                out_buf = pb.tensor[: len(members) * 84 * args.t_out * 120 * 240].view(len(members), 84, args.t_out, 120, 240)
        except Exception:
            out_buf = None
    run = Rollout(args.strong_model, members, args, dev, ae=ae, model=model, out=out_buf)
    fields, ms, host_ms, launches, _ = timed_loop(run, steps, warmup, barrier, world, dev, lib)
    step_ms = ms / steps
    value = ens_total * args.t_out * steps / (ms * 1e-3)
    res = {"model": f"ladcast_{args.strong_model}", "ensemble_total": ens_total, "members_per_rank": per_rank,
           "scaling": "strong", "value": value, "unit": UNIT, "ms_per_step": step_ms, "steps": steps, "warmup": warmup,
           "gpu_launches_per_step": launches // max(1, steps),
           "balance_bound": round(ens_total / (world * max(per_rank)), 4),
           "step_algorithmic_tflops_per_gpu": round(flops_per_member_step(args.strong_model, args.t_out, args.denoise_steps)
                                                    * max(per_rank) * args.t_out / (step_ms * 1e-3) / 1e12, 1)}
    probe = run.probe_host(lib)
    res["launch_probe"] = probe
    if probe is not None:
        ratio = probe["host_enqueue_us_per_denoiser_call"] / max(1e-9, probe["device_us_per_denoiser_call"])
        res["limiter"] = (f"host launch rate: enqueueing a denoiser call takes {ratio:.2f}x its device time" if ratio > 0.8 else
                          f"device (the rank with the most members): enqueueing a denoiser call costs {ratio:.2f}x its device "
                          "time, so the launch queue stays ahead of the GPU")
    # ---- metrics over the sharded ensemble (config 5): exchange + kernel, timed; checked against one GPU
    g = torch.Generator("cpu").manual_seed(11)
    truth = torch.randn((84, args.t_out, 120, 240), generator=g).to(dev)
    if fields is None:
        fields = torch.empty((0, 84, args.t_out, 120, 240), device=dev)
    if world > 1:
        tm = {}
        for _ in range(2):  # first pass opens the NCCL peer connections
            tabs = ensemble_metrics_distributed(fields, truth, timings=tm)
        m_max = max(per_rank)
        pad = torch.zeros((m_max,) + tuple(fields.shape[1:]), device=dev)
        pad[: fields.shape[0]] = fields
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad)
        full = torch.cat([bufs[q][: per_rank[q]] for q in range(world)], dim=0)
        del bufs, pad
        want = ensemble_metrics(full, truth)
        ok = all(torch.allclose(tabs[k], want[k], rtol=1e-9, atol=1e-12, equal_nan=True) for k in want)
        t = torch.tensor([tm["exchange_ms"], tm["kernel_ms"], float(tm["bytes_sent"]), 0.0 if ok else 1.0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ex_ms, k_ms, by, bad = [float(v) for v in t.tolist()]
        res["metrics"] = {"exchange_ms": round(ex_ms, 3), "kernel_ms": round(k_ms, 3), "bytes": int(by),
                          "exchange_gbs_per_gpu": round(by / (ex_ms * 1e-3) / 1e9, 1) if ex_ms > 0 else None,
                          "matches_single_gpu": bad == 0.0, "rtol": 1e-9,
                          "what": "ensemble_metrics_distributed on the last AR step's decoded fields (max over ranks; "
                                  "bytes = fp32 bytes one rank sends) vs ensemble_metrics on the all-gathered fields"}
        del full
        # the same metrics with NO exchange step: peer-memory reads inside the reduction kernel (exchange="p2p")
        try:
            tp = {}
            for _ in range(3):  # first pass allocates + maps the peer-visible buffers (CUDA IPC)
                tabs_p = ensemble_metrics_distributed(fields, truth, timings=tp, exchange="p2p")
            okp = all(torch.allclose(tabs_p[k], want[k], rtol=1e-9, atol=1e-12, equal_nan=True) for k in want)
            t = torch.tensor([tp["kernel_ms"], 0.0 if okp else 1.0, tp["exchange_ms"]], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            res["metrics"]["p2p_fused"] = {
                "kernel_ms": round(float(t[0]), 3), "stage_and_barrier_ms": round(float(t[2]), 3),
                "fields_in_peer_buffer": out_buf is not None, "matches_single_gpu": float(t[1]) == 0.0,
                "remote_gbs_per_gpu": round(by / (float(t[0]) * 1e-3) / 1e9, 1) if float(t[0]) > 0 else None,
                "what": "lc_metrics_accumulate_ptrs: every rank reduces its plane slice reading the other ranks' members "
                        "in place over NVLink (CUDA-IPC peer memory) - exchange and reduction are one kernel"}
        except Exception as ex:  # never lose the bench line to the optional path
            res["metrics"]["p2p_fused"] = {"error": repr(ex)[:300]}
    else:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ensemble_metrics(fields, truth)
        e0.record()
        tabs = ensemble_metrics(fields, truth)
        e1.record()
        torch.cuda.synchronize()
        res["metrics"] = {"exchange_ms": 0.0, "kernel_ms": round(e0.elapsed_time(e1), 3), "bytes": 0,
                          "matches_single_gpu": True, "what": "single GPU: no exchange"}
    res["finite"] = bool(torch.isfinite(fields).all().item()) and all(bool(torch.isfinite(v).all()) for v in tabs.values())
    del run, fields
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="375M", choices=list(MODEL_CFG))
    ap.add_argument("--ens", type=int, default=20, help="ensemble members per GPU (weak-scaling headline)")
    ap.add_argument("--denoise-steps", type=int, default=20)
    ap.add_argument("--t-out", type=int, default=4)
    ap.add_argument("--strong-model", default="1.6B", choices=list(MODEL_CFG))
    ap.add_argument("--strong-ens", default="20,50", help="fixed ensemble sizes of the strong-scaling leg (configs 5, 4)")
    ap.add_argument("--no-strong", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-metrics", action="store_true")
    ap.add_argument("--no-roofline", action="store_true", help="skip the extra event-profiled step (for ncu launch lists)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3

    import torch.distributed as dist

    from ladcast_b200 import _lib
    from ladcast_b200.pipelines.utils import roll_out_latent

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass

    # ---- headline: weak scaling, `--ens` members per GPU (global member ids so that ranks draw distinct noise)
    members = list(range(rank * args.ens, (rank + 1) * args.ens))
    run = Rollout(args.model, members, args, dev)
    out, ms, host_ms, launches, clk = timed_loop(run, args.steps, args.warmup, barrier, world, dev, lib,
                                                 clock_index=local if rank == 0 else None)
    finite = bool(torch.isfinite(out).all().item())
    units = args.ens * args.t_out * args.steps * world
    value = units / (ms * 1e-3)
    step_ms = ms / args.steps
    launch_probe = run.probe_host(lib)

    # ---- rooflines: one more identical step with per-launch CUDA events
    roofline = None
    if not args.no_roofline:
        per_class = class_rooflines(profiled_step(run, lib, _lib), peaks, step_ms)
        roofline = dict(per_class.get("gemm_tc", {}))
        # DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum per launch) comes from ncu, never from this run:
        # profiles/r02_*.md hold the per-kernel captures; no constant is pasted here
        roofline["traffic"] = None
        roofline["traffic_source"] = "profiles/ (ncu --set full captures, per kernel); not measurable inside bench.py"
        roofline["secondary"] = [dict(v, **{"class": k}) for k, v in per_class.items() if k != "gemm_tc"]
        roofline["step_algorithmic_tflops"] = round(flops_per_member_step(args.model, args.t_out, args.denoise_steps)
                                                    * args.ens * args.t_out / (step_ms * 1e-3) / 1e12, 1)

    # ---- the metrics kernel on the decoded fields
    metrics = None
    if not args.no_metrics and rank == 0:
        metrics = metrics_leg(out, lib, _lib, peaks)

    # ---- end to end through the public API with host buffers (H2D of inputs, D2H of decoded fields, each step)
    e2e = None
    if not args.no_e2e:
        k_e2e = min(args.steps, 5)
        host_known = run.known0.clone().pin_memory()
        host_out = torch.empty((k_e2e, args.ens, 84, args.t_out, 120, 240), dtype=torch.float32, pin_memory=True)
        barrier()
        t0 = time.perf_counter()
        roll_out_latent(run.pipe, run.ae, host_known, 2018010100, args.ens, run.lat_mean, run.lat_std, run.fld_mean,
                        run.fld_std, num_inference_steps=args.denoise_steps, return_seq_len=args.t_out,
                        sampler_type="pipeline", member_indices=members, out=host_out, max_ar_steps=k_e2e)
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        noise_bytes = args.ens * 84 * args.t_out * 15 * 30 * 4
        e2e = {"value": args.ens * args.t_out * k_e2e * world / dt, "unit": UNIT,
               "h2d_bytes_per_step": noise_bytes + 84 * 15 * 30 * 4,
               "d2h_bytes_per_step": args.ens * 84 * args.t_out * 120 * 240 * 4, "steps": k_e2e,
               "api": "ladcast_b200.pipelines.utils.roll_out_latent (pinned host in/out, async D2H per AR step)"}
        del host_out

    # ---- strong scaling of the north-star configuration (1.6B, fixed ensembles sharded over the ranks)
    strong = None
    ae = run.ae
    run.release()
    del out
    torch.cuda.empty_cache()
    if not args.no_strong:
        from ladcast_b200.models import LaDCastTransformer3DModel

        strong = []
        torch.manual_seed(1234)
        big = LaDCastTransformer3DModel.from_config(denoiser_kwargs(args.strong_model)).to(dev)  # built once, both ensembles
        for i, e in enumerate(int(v) for v in args.strong_ens.split(",") if v):
            k = min(args.steps, 3 if i == 0 else 2)
            strong.append(strong_leg(args, e, rank, world, dev, ae, big, barrier, lib, _lib, steps=k, warmup=3 if i == 0 else 2))
        big._release()

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_base, _ = cpu_reference_sample(args.model, args.t_out, args.denoise_steps)
        try:
            cpu_base["config1"] = cpu_config1()
        except Exception as ex:  # the baseline is a report, never a reason to lose the GPU line
            cpu_base["config1"] = {"error": repr(ex)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args, world),
                "roofline": roofline, "cpu_baseline": cpu_base, "e2e": e2e, "gpu_launches": int(launches),
                "launch_probe": launch_probe, "metrics": metrics, "strong": strong,
                "clocks": clk, "finite": finite, "impl": "ours"}
        print(json.dumps(line))
    if world > 1:
        try:
            from ladcast_b200.evaluate.utils import release_peer_buffers

            release_peer_buffers()
        except Exception:
            pass
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
