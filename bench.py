#!/usr/bin/env python
"""Benchmark of the LaDCast ensemble-rollout hot path (BASELINE.json metric: ensemble-member 6 h-steps / second).

  python bench.py --gpus N --steps K --warmup W            # ours: sm_100a CUDA path, one process per GPU
  python bench.py --impl reference --gpus N ...             # the reference algorithm (CPU oracle port) on host cores

One "step" = one autoregressive step of the rollout for the members resident on a GPU: `num_inference_steps`
denoiser calls of the DPM-Solver++ sampler (T_out lead steps at once), the scheduler updates, feeding the last frame
back, latent de-normalisation and the DC-AE decode of every (member, lead) frame.  It yields ens * T_out
member-6h-steps.  Members shard across GPUs with no communication inside the rollout (weak scaling: every rank
runs `--ens` members).  Prints ONE JSON line on rank 0.
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


def cpu_reference_sample(model_name, t_out, n_denoise, repeats=1):
    """The reference algorithm (oracle/ladcast_oracle.py, a CPU fp32 port validated against the unmodified reference)
    on the host cores: a bounded sample of the same workload — 2 denoiser calls (1 member, T_out lead steps, 2250
    tokens) + 1 decoded frame — scaled to one member's AR step (n_denoise calls + T_out frames)."""
    from oracle import ladcast_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.denoiser_config(model_name)
    sd = O.make_state_dict(O.denoiser_param_shapes(cfg), 1)
    acfg = O.dcae_config()
    asd = O.make_state_dict(O.dcae_decoder_param_shapes(acfg), 2)
    g = torch.Generator("cpu").manual_seed(0)
    x = torch.randn((1, 84, t_out, 15, 30), generator=g)
    cond = torch.randn((1, 84, 1, 15, 30), generator=g) * 0.5
    z = torch.randn((1, 84, 15, 30), generator=g)
    ts = torch.tensor([2018010100])
    vals = []
    with torch.no_grad():
        O.denoiser_forward(sd, cfg, x, torch.tensor([0.5]), cond, ts)  # warm-up
        for _ in range(repeats):
            t0 = time.perf_counter()
            for tt in (0.7, -0.3):
                O.denoiser_forward(sd, cfg, x, torch.tensor([tt]), cond, ts)
            t_call = (time.perf_counter() - t0) / 2
            t0 = time.perf_counter()
            O.dcae_decode(asd, acfg, z)
            t_frame = time.perf_counter() - t0
            vals.append((t_out / (n_denoise * t_call + t_out * t_frame), t_call, t_frame))
    v = sum(a for a, _, _ in vals) / len(vals)
    return {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"oracle port (CPU fp32, torch {torch.__version__}, {cores} threads): 2 denoiser calls B=1 T_out={t_out} "
                      f"({vals[-1][1]:.2f} s/call) + 1 decoded frame ({vals[-1][2]:.2f} s), scaled to {n_denoise} calls + "
                      f"{t_out} frames per member AR step"}, vals


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="375M", choices=list(MODEL_CFG))
    ap.add_argument("--ens", type=int, default=20, help="ensemble members per GPU")
    ap.add_argument("--denoise-steps", type=int, default=20)
    ap.add_argument("--t-out", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-roofline", action="store_true", help="skip the extra event-profiled step (for ncu launch lists)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3

    import torch.distributed as dist

    from ladcast_b200 import _lib
    from ladcast_b200.models import AutoencoderDC, LaDCastTransformer3DModel
    from ladcast_b200.pipelines import AutoRegressive2DPipeline, EDMDPMSolverMultistepScheduler
    from ladcast_b200.pipelines.utils import decode_latent_ens, ensemble_AR_sampler, roll_out_latent

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    torch.manual_seed(1234)  # identical random-init weights on every rank
    model = LaDCastTransformer3DModel.from_config(denoiser_kwargs(args.model)).to(dev)
    ae = AutoencoderDC(**DCAE_KW).to(dev)
    pipe = AutoRegressive2DPipeline(model, EDMDPMSolverMultistepScheduler())
    members = list(range(rank * args.ens, (rank + 1) * args.ens))  # global member ids of this rank
    g = torch.Generator("cpu").manual_seed(7)
    known0 = torch.randn((1, 84, 1, 15, 30), generator=g) * 0.5
    lat_mean, lat_std = torch.randn(84, generator=g) * 0.1, torch.rand(84, generator=g) + 0.5
    fld_mean, fld_std = torch.randn(84, generator=g), torch.rand(84, generator=g) + 0.5
    lm, ls = lat_mean.to(dev)[None, :, None, None, None], lat_std.to(dev)[None, :, None, None, None]
    fm, fs = fld_mean.to(dev), fld_std.to(dev)
    state = {"known": known0.to(dev), "step": 0}

    def ar_step():
        stamp = torch.tensor([2018010100 + 0])  # date embedding is recomputed every AR step like the reference
        s = ensemble_AR_sampler(pipe, sample_size=args.ens, return_seq_len=args.t_out,
                                num_inference_steps=args.denoise_steps, known_latents=state["known"], timestamps=stamp,
                                sampler_type="pipeline", device=dev, member_indices=members)
        state["known"] = s[:, :, -1:].clone()
        phys = (s / 0.5) * ls + lm
        fields = decode_latent_ens(ae, phys, fm, fs)
        state["step"] += 1
        return fields

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        out = ar_step()
    barrier()
    clocks = ClockSampler(local) if rank == 0 else None
    launches0 = lib.lc_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = ar_step()
    e1.record()
    barrier()
    launches = lib.lc_launch_count() - launches0
    clk = clocks.stop() if clocks else None
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    finite = bool(torch.isfinite(out).all().item())
    units = args.ens * args.t_out * args.steps * world
    value = units / (ms * 1e-3)

    # ---- roofline of the dominant kernel (tcgen05 GEMM): one more identical step with per-launch CUDA events
    import ctypes

    lib.lc_prof_enable(1)
    if not args.no_roofline:
        ar_step()
    torch.cuda.synchronize()
    pms, pfl, pln = (ctypes.c_double * 3)(), (ctypes.c_double * 3)(), (ctypes.c_longlong * 3)()
    _lib.check(lib.lc_prof_collect(pms, pfl, pln), "lc_prof_collect")
    lib.lc_prof_enable(0)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks else "fallback 1.4 PFLOP/s sustained"
    classes = {}
    for i, nme in enumerate(("gemm_tc", "attention_tc", "sphere_conv_tc")):
        if pln[i]:
            classes[nme] = {"launches": int(pln[i]), "ms": round(pms[i], 3), "tflops": round(pfl[i] / (pms[i] * 1e-3) / 1e12, 1)}
    gemm_ach = pfl[0] / (pms[0] * 1e-3) / 1e12 if pms[0] > 0 else 0.0
    step_ms = ms / args.steps
    roofline = {"bound": "tensor", "kernel": "gemm_tc2_kernel / gemm_tc_kernel (tcgen05 bf16 GEMM, CTA-pair; denoiser linears + decoder 1x1)",
                "achieved": round(gemm_ach, 1), "peak": peak_tf, "unit": "TFLOP/s", "frac": round(gemm_ach / peak_tf, 4),
                "peak_source": peak_src,
                # dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged over 36 consecutive launches of this
                # kernel in a 375M B=20 denoiser call (profiles/r01_gemm_ncu_final.md); only valid for that workload
                "traffic": 383.1e6 if (args.model == "375M" and args.ens == 20 and args.t_out == 4) else None,
                "traffic_source": "profiles/r01_gemm_ncu_final.md (ncu, per-launch average, bytes)",
                "avg_launch_ms": round(pms[0] / max(1, pln[0]), 4), "share_of_step": round(pms[0] / step_ms, 3),
                "classes": classes,
                "step_algorithmic_tflops": round(flops_per_member_step(args.model, args.t_out, args.denoise_steps) * args.ens
                                                 * args.t_out / (step_ms * 1e-3) / 1e12, 1)}

    # ---- end to end through the public API with host buffers (H2D of inputs, D2H of decoded fields, each step)
    e2e = None
    if not args.no_e2e:
        k_e2e = min(args.steps, 5)
        host_known = known0.clone().pin_memory()
        host_out = torch.empty((k_e2e, args.ens, 84, args.t_out, 120, 240), dtype=torch.float32, pin_memory=True)
        barrier()
        t0 = time.perf_counter()
        roll_out_latent(pipe, ae, host_known, 2018010100, args.ens, lat_mean, lat_std, fld_mean, fld_std,
                        num_inference_steps=args.denoise_steps, return_seq_len=args.t_out, sampler_type="pipeline",
                        member_indices=members, out=host_out, max_ar_steps=k_e2e)
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        noise_bytes = args.ens * 84 * args.t_out * 15 * 30 * 4
        e2e = {"value": args.ens * args.t_out * k_e2e * world / dt, "unit": UNIT,
               "h2d_bytes_per_step": noise_bytes + (84 * 15 * 30 * 4 if True else 0),
               "d2h_bytes_per_step": args.ens * 84 * args.t_out * 120 * 240 * 4, "steps": k_e2e,
               "api": "ladcast_b200.pipelines.utils.roll_out_latent (pinned host in/out, async D2H per AR step)"}

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_base, _ = cpu_reference_sample(args.model, args.t_out, args.denoise_steps)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args, world),
                "roofline": roofline, "cpu_baseline": cpu_base, "e2e": e2e, "gpu_launches": int(launches),
                "clocks": clk, "finite": finite, "impl": "ours"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
