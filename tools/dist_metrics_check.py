"""torchrun check (N GPUs): member-sharded metrics over NCCL all-to-all == single-GPU metrics on all members."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from ladcast_b200.evaluate.utils import ensemble_metrics, ensemble_metrics_distributed
from ladcast_b200.pipelines.utils import member_shard

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
g = torch.Generator("cpu").manual_seed(5)
M = 5 * world - 1  # uneven shards
fields = torch.randn((M, 84, 4, 120, 240), generator=g)
truth = torch.randn((84, 4, 120, 240), generator=g)
truth[82, :, :10] = float("nan")
mine = list(member_shard(M, rank, world))
tabs = ensemble_metrics_distributed(fields[mine].cuda().contiguous(), truth.cuda())
want = ensemble_metrics(fields.cuda(), truth.cuda())
ok = all(torch.allclose(tabs[k], want[k], rtol=1e-9, atol=1e-12, equal_nan=True) for k in want)
print(f"rank {rank}/{world}: members {mine[0]}..{mine[-1]} distributed metrics match single-GPU: {ok}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
