"""torchrun check (N GPUs): member-sharded metrics == single-GPU metrics on all members, for both exchange modes —
"nccl" (grouped send/recv member->plane re-shard, then the kernel) and "p2p" (no exchange: the kernel reads the other
ranks' members in place over NVLink) — with CUDA-event timings.  Rank 0 appends one JSON line to
gpurun_out/dist_metrics.jsonl."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from ladcast_b200.evaluate.utils import ensemble_metrics, ensemble_metrics_distributed, release_peer_buffers
from ladcast_b200.pipelines.utils import member_shard

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
g = torch.Generator("cpu").manual_seed(5)
res, all_ok = {"world": world, "cases": []}, True
for M in (5 * world - 1, 20, 50):  # uneven shards; BASELINE configs 5 and 4
    fields = torch.randn((M, 84, 4, 120, 240), generator=g)
    truth = torch.randn((84, 4, 120, 240), generator=g)
    truth[82, :, :10] = float("nan")
    mine = list(member_shard(M, rank, world))
    local_f = fields[mine].cuda().contiguous()
    want = ensemble_metrics(fields.cuda(), truth.cuda())
    case = {"members": M, "members_per_rank": [len(member_shard(M, r, world)) for r in range(world)]}
    for mode in ("nccl", "p2p"):
        tm = {}
        for _ in range(3):  # first pass: NCCL connections / IPC mapping
            tabs = ensemble_metrics_distributed(local_f, truth.cuda(), timings=tm, exchange=mode)
        ok = all(torch.allclose(tabs[k], want[k], rtol=1e-9, atol=1e-12, equal_nan=True) for k in want)
        t = torch.tensor([tm["exchange_ms"], tm["kernel_ms"], 0.0 if ok else 1.0], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ex, k, bad = [float(v) for v in t.tolist()]
        case[mode] = {"exchange_ms": round(ex, 3), "kernel_ms": round(k, 3), "total_ms": round(ex + k, 3),
                      "bytes_sent_per_rank": int(tm["bytes_sent"]), "matches_single_gpu": bad == 0.0}
        all_ok = all_ok and bad == 0.0
        print(f"rank {rank}/{world}: M={M} {mode}: match={ok} exchange {tm['exchange_ms']:.3f} ms kernel {tm['kernel_ms']:.3f} ms", flush=True)
    res["cases"].append(case)
    del fields, local_f
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/dist_metrics.jsonl", "a") as f:
        f.write(json.dumps(res) + "\n")
    print(json.dumps(res))
release_peer_buffers()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if all_ok else 1)
