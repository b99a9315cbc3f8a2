import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import ladcast_oracle as O
from ladcast_b200.models import LaDCastTransformer3DModel
name, B = (sys.argv[1] if len(sys.argv) > 1 else "375M"), int(sys.argv[2]) if len(sys.argv) > 2 else 20
m = LaDCastTransformer3DModel.from_config(O.denoiser_config(name)).to("cuda")
x = torch.randn(B, 84, 4, 15, 30, device="cuda"); cond = torch.randn(B, 84, 1, 15, 30, device="cuda")
t = torch.full((B,), 0.5, device="cuda"); ts = torch.tensor([2018010100])
with m.cached_conditioning(cond, ts, t_out=4):
    for _ in range(4):
        m(x, t, cond, time_elapsed=ts)
torch.cuda.synchronize()
