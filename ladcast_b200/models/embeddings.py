"""Host-side tables of the denoiser: RoPE cos/sin grids and the year-progress embedding.

Mirrors (own implementation) `LaDCastRotaryPosEmbed_from_grid.forward` (reference models/embeddings.py:274-327, fed by
LaDCast_3D_model.py:885-938) and `get_year_sincos_embedding` (models/embeddings.py:428-520).  Both are tiny,
step-invariant and computed once per geometry / per AR step on the host, then uploaded."""
import math
from datetime import datetime

import numpy as np
import torch


def _axis_table(dim, pos, theta):
    # diffusers get_1d_rotary_pos_embed(use_real=True): one angle per feature pair, repeated for both members
    inv = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float32)[: dim // 2] / dim))
    ang = torch.outer(pos.to(torch.float32), inv)
    return ang.cos().repeat_interleave(2, dim=1).float(), ang.sin().repeat_interleave(2, dim=1).float()


def rope_tables(config, t_in, t_out, height, width):
    """Returns ((cos_pred, sin_pred), (cos_cond, sin_cond)), each [T*H*W, head_dim] fp32 on CPU.
    Token order t-major, then latitude, then longitude ('ij' meshgrid); temporal coordinate of cond frames is
    -T_in+1..0 and of pred frames 1..T_out; the spatial grid is linspace(start, end) (deg2rad'ed if configured)."""
    start, end = config["rope_spatial_grid_start_pos"], config["rope_spatial_grid_end_pos"]
    if config.get("spatial_deg2rad", False):
        start = [float(np.deg2rad(v)) for v in start]
        end = [float(np.deg2rad(v)) for v in end]
    lat = torch.linspace(start[0], end[0], steps=height, dtype=torch.float32)
    lon = torch.linspace(start[1], end[1], steps=width, dtype=torch.float32)
    theta = config["rope_theta"]

    def build(tcoord, dims):
        tt, la, lo = torch.meshgrid(tcoord, lat, lon, indexing="ij")
        parts = [_axis_table(dims[0], tt.reshape(-1), theta), _axis_table(dims[1], la.reshape(-1), theta),
                 _axis_table(dims[2], lo.reshape(-1), theta)]
        return (torch.cat([p[0] for p in parts], dim=1).contiguous(), torch.cat([p[1] for p in parts], dim=1).contiguous())

    pred = build(torch.arange(1, t_out + 1, dtype=torch.float32), config["rope_axes_dim"])
    cond = build(torch.arange(-t_in + 1, 1, dtype=torch.float32), config["conditioning_tensor_rope_axes_dim"])
    return pred, cond


def year_fraction(stamp):
    """YYYYMMDDHH -> elapsed fraction of that (leap-aware) calendar year."""
    s = str(int(stamp))
    now = datetime(int(s[:4]), int(s[4:6]), int(s[6:8]), int(s[8:10]))
    a, b = datetime(now.year, 1, 1), datetime(now.year + 1, 1, 1)
    return (now - a).total_seconds() / (b - a).total_seconds()


def year_sincos_embedding(stamps, embedding_dim=256, max_period=10000):
    """[n, embedding_dim] fp32: sin(2*pi*p*k)*m_k | cos(2*pi*p*k)*m_k, k = 1..half, m_k = exp(-ln(max_period)(k-1)/half)."""
    half = embedding_dim // 2
    p = torch.tensor([year_fraction(v) for v in stamps], dtype=torch.float32)
    k = torch.arange(1, half + 1).float()
    mag = torch.exp(-math.log(max_period) * torch.arange(0, half).float() / half)
    arg = (2 * math.pi * p.reshape(-1, 1)) * k.reshape(1, -1)
    return torch.cat([torch.sin(arg) * mag[None], torch.cos(arg) * mag[None]], dim=1).contiguous()
