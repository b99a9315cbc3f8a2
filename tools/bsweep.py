"""Member-steps/s of one AR step (20 denoiser calls + decode) versus the number of members resident on the GPU, for both
models — the small-batch end is what the 8-GPU strong-scaling shards see (20 members / 8 GPUs = 2-3 per GPU).
  python tools/bsweep.py [375M,1.6B] [2,3,7,13,20]      -> one JSON line per (model, B) + gpurun_out/bsweep.json"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from ladcast_b200 import _lib

models = (sys.argv[1] if len(sys.argv) > 1 else "375M,1.6B").split(",")
Bs = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "2,3,7,13,20").split(",")]
args = argparse.Namespace(denoise_steps=20, t_out=4)
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
lib = _lib.load()
res, ae = [], None
for name in models:
    for B in Bs:
        run = bench.Rollout(name, list(range(B)), args, dev, ae=ae)
        ae = run.ae
        _, ms, host_ms, launches, _ = bench.timed_loop(run, 2, 2, torch.cuda.synchronize, 1, dev, lib)
        cls = bench.profiled_step(run, lib, _lib)
        r = {"model": name, "members": B, "ms_per_step": ms / 2, "member_steps_per_s": B * 4 * 2 / (ms * 1e-3),
             "per_member_ms": ms / 2 / B, "launch_probe": run.probe_host(lib), "launches_per_step": launches // 2,
             "tflops": {k: round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 1) for k, v in cls.items() if v["flops"] > 0},
             "class_ms": {k: round(v["ms"], 2) for k, v in cls.items()}}
        print(json.dumps(r), flush=True)
        res.append(r)
        run.release()
        del run
        torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/bsweep.json", "w"), indent=1)
