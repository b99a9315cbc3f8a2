import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ladcast_b200.models import AutoencoderDC
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import DCAE_KW
n = int(sys.argv[1]) if len(sys.argv) > 1 else 80
ae = AutoencoderDC(**DCAE_KW).to("cuda")
z = torch.randn(n, 84, 15, 30, device="cuda")
mean, std = torch.zeros(84), torch.ones(84)
for _ in range(2):
    ae.decode_fused(z, mean, std)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    ae.decode_fused(z, mean, std)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"decode {n} frames: {ms:.2f} ms  ({n*0.7814/ms:.1f} TF/s algorithmic, {ms/n:.3f} ms/frame)")
