"""TEST INFRASTRUCTURE — CPU restatement (plain PyTorch fp32/fp64, functional style over a state-dict) of the
LaDCast ensemble-rollout hot path.  It is the checker for the CUDA path, never the product: only tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs may import it.

Every function cites the reference file:line it follows (paths relative to /root/reference/ladcast).  The
arithmetic that lives in the un-vendored third-party `diffusers==0.32.1` (pyproject.toml:34) is restated from
its published source; for those pieces PARITY IS UNPINNED by any reference-owned test.
What *is* pinned: `tests/test_oracle_golden.py` checks this file against vectors produced by running the
UNMODIFIED reference modules (with oracle/shim standing in for diffusers) via oracle/make_golden.py, and against
the only known-answer vector in the reference (models/sphere_conv.py:141-172).
"""
from __future__ import annotations

import math
from datetime import datetime
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

# --------------------------------------------------------------------------------------------------------------
# configs (configs/ladcast_375M.yaml:1-31, ladcast_1.6B.yaml:5-9, DC_AE_84_ft.yaml:1-48)
# --------------------------------------------------------------------------------------------------------------


def denoiser_config(name: str = "375M", **over) -> dict:
    cfg = dict(
        in_channels=84, out_channels=84, num_attention_heads=12, attention_head_dim=128, num_layers=2,
        num_single_layers=4, num_refiner_layers=1, mlp_ratio=4, patch_size=1, patch_size_t=1, qk_norm="rms_norm",
        rope_theta=256.0, rope_axes_dim=[16, 56, 56], rope_spatial_grid_start_pos=[-499.5, 5.25],
        rope_spatial_grid_end_pos=[508.5, 353.25], spatial_deg2rad=True, conditioning_tensor_in_channels=84,
        conditioning_tensor_rope_axes_dim=[16, 56, 56], incl_time_elapsed=True,
    )
    if name == "1.6B":
        cfg.update(num_attention_heads=16, num_layers=5, num_single_layers=10, num_refiner_layers=3)
    elif name == "tiny":  # test-only geometry: same head_dim/rope split, 2 heads, one block of each kind
        cfg.update(num_attention_heads=2, num_layers=1, num_single_layers=1, num_refiner_layers=1)
    elif name != "375M":
        raise ValueError(name)
    cfg.update(over)
    return cfg


def dcae_config(name: str = "V0.1.X", **over) -> dict:
    cfg = dict(
        in_channels=89, out_channels=89, latent_channels=84, attention_head_dim=32,
        decoder_block_types=["ResBlock", "ResBlock", "EfficientViTBlock", "EfficientViTBlock"],
        decoder_block_out_channels=[252, 504, 504, 1008], decoder_layers_per_block=[4, 4, 4, 4],
        decoder_qkv_multiscales=[[], [], [5], [5]], static_channels=5,
    )
    if name == "tiny":  # test-only: same topology, 1 layer per stage, narrow channels
        cfg.update(decoder_block_out_channels=[84, 168, 168, 336], decoder_layers_per_block=[1, 1, 1, 1])
    elif name != "V0.1.X":
        raise ValueError(name)
    # the shipped config mirrors the decoder in the encoder (configs/DC_AE_84_ft.yaml)
    for k in ("block_types", "block_out_channels", "layers_per_block", "qkv_multiscales"):
        cfg["encoder_" + k] = list(cfg["decoder_" + k])
    cfg.update(over)
    return cfg


# --------------------------------------------------------------------------------------------------------------
# deterministic weights: both the golden generator (which loads them into the unmodified reference modules) and
# the GPU-side tests build identical state-dicts from (key, shape) alone, so no weight file has to be committed.
# --------------------------------------------------------------------------------------------------------------


def _key_seed(key: str, salt: int) -> int:
    h = 1469598103934665603
    for ch in (key + f"#{salt}").encode():
        h = ((h ^ ch) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h & 0x7FFFFFFF


def det_tensor(key: str, shape: Sequence[int], salt: int = 0) -> torch.Tensor:
    """Deterministic pseudo-random parameter.  Linear/conv weights ~ U(-1,1)/sqrt(fan_in) * 1.7 (unit-ish gain),
    norm scales ~ 1 + 0.1 N(0,1), biases ~ 0.1 N(0,1), so every term of the network is exercised."""
    g = torch.Generator("cpu").manual_seed(_key_seed(key, salt))
    shape = tuple(int(s) for s in shape)
    if key.endswith(".weight") and len(shape) >= 2:
        fan_in = int(np.prod(shape[1:]))
        return (torch.rand(shape, generator=g) * 2 - 1) * (1.7 / math.sqrt(fan_in))
    if key.endswith(".weight"):
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    return 0.1 * torch.randn(shape, generator=g)


def denoiser_param_shapes(cfg: dict) -> Dict[str, Tuple[int, ...]]:
    """State-dict keys/shapes of LaDCastTransformer3DModel (models/LaDCast_3D_model.py:624-766)."""
    d = cfg["num_attention_heads"] * cfg["attention_head_dim"]
    hd = cfg["attention_head_dim"]
    C, Cc, Co = cfg["in_channels"], cfg["conditioning_tensor_in_channels"], cfg["out_channels"]
    mlp = int(d * cfg["mlp_ratio"])
    s: Dict[str, Tuple[int, ...]] = {}

    def lin(name, o, i):
        s[name + ".weight"] = (o, i)
        s[name + ".bias"] = (o,)

    s["x_embedder.proj.weight"] = (d, C, 1, 1, 1)
    s["x_embedder.proj.bias"] = (d,)
    s["context_embedder.proj.weight"] = (d, Cc, 1, 1, 1)
    s["context_embedder.proj.bias"] = (d,)
    for pre in ("context_refiner.time_text_embed", "time_text_embed"):
        lin(pre + ".timestep_embedder.linear_1", d, 256)
        lin(pre + ".timestep_embedder.linear_2", d, d)
        lin(pre + ".text_embedder.linear_1", d, d)
        lin(pre + ".text_embedder.linear_2", d, d)
    lin("context_refiner.proj_in", d, d)
    for i in range(cfg["num_refiner_layers"]):
        p = f"context_refiner.token_refiner.refiner_blocks.{i}"
        for n in ("norm1", "norm2"):
            s[f"{p}.{n}.weight"] = (d,)
            s[f"{p}.{n}.bias"] = (d,)
        for n in ("to_q", "to_k", "to_v"):
            lin(f"{p}.attn.{n}", d, d)
        s[f"{p}.attn.norm_q.weight"] = (hd,)
        s[f"{p}.attn.norm_k.weight"] = (hd,)
        lin(f"{p}.ff.net.0.proj", mlp, d)
        lin(f"{p}.ff.net.2", d, mlp)
        lin(f"{p}.norm_out.linear", 2 * d, d)
    if cfg.get("incl_time_elapsed", False):
        lin("time_elapsed_embed.linear_1", 2 * d, 256)
        lin("time_elapsed_embed.linear_2", 2 * d, 2 * d)
    for i in range(cfg["num_layers"]):
        p = f"transformer_blocks.{i}"
        lin(f"{p}.norm1.linear", 6 * d, d)
        lin(f"{p}.norm1_context.linear", 6 * d, d)
        for n in ("to_q", "to_k", "to_v", "add_k_proj", "add_v_proj", "add_q_proj"):
            lin(f"{p}.attn.{n}", d, d)
        for n in ("norm_q", "norm_k", "norm_added_q", "norm_added_k"):
            s[f"{p}.attn.{n}.weight"] = (hd,)
        lin(f"{p}.attn.to_out.0", d, d)
        lin(f"{p}.attn.to_add_out", d, d)
        for ff in ("ff", "ff_context"):
            lin(f"{p}.{ff}.net.0.proj", mlp, d)
            lin(f"{p}.{ff}.net.2", d, mlp)
    for i in range(cfg["num_single_layers"]):
        p = f"single_transformer_blocks.{i}"
        for n in ("to_q", "to_k", "to_v"):
            lin(f"{p}.attn.{n}", d, d)
        s[f"{p}.attn.norm_q.weight"] = (hd,)
        s[f"{p}.attn.norm_k.weight"] = (hd,)
        lin(f"{p}.norm.linear", 3 * d, d)
        lin(f"{p}.proj_mlp", mlp, d)
        lin(f"{p}.proj_out", d, d + mlp)
    lin("norm_out.linear", 2 * d, d)
    lin("proj_out", Co, d)
    return s


def dcae_decoder_layout(cfg: dict) -> List[Tuple[str, str, int, int]]:
    """Ordered decoder.up_blocks: (kind, key-prefix, C_in, C_out).  models/DCAE.py:669-694."""
    chans = cfg["decoder_block_out_channels"]
    layers = cfg["decoder_layers_per_block"]
    types = cfg["decoder_block_types"]
    out: List[Tuple[str, str, int, int]] = []
    j = 0
    n = len(chans)
    for i in reversed(range(n)):
        if i < n - 1 and layers[i] > 0:
            out.append(("up", f"decoder.up_blocks.{j}", chans[i + 1], chans[i]))
            j += 1
        for _ in range(layers[i]):
            kind = "res" if types[i] == "ResBlock" else "evit"
            out.append((kind, f"decoder.up_blocks.{j}", chans[i], chans[i]))
            j += 1
    return out


def dcae_decoder_param_shapes(cfg: dict) -> Dict[str, Tuple[int, ...]]:
    """State-dict keys/shapes of AutoencoderDC.decoder (models/DCAE.py:634-715)."""
    s: Dict[str, Tuple[int, ...]] = {}
    chans = cfg["decoder_block_out_channels"]
    hd = cfg["attention_head_dim"]
    s["decoder.conv_in.weight"] = (chans[-1], cfg["latent_channels"], 3, 3)
    s["decoder.conv_in.bias"] = (chans[-1],)
    for kind, p, ci, co in dcae_decoder_layout(cfg):
        if kind == "up":
            s[f"{p}.conv.weight"] = (4 * co, ci, 3, 3)
            s[f"{p}.conv.bias"] = (4 * co,)
        elif kind == "res":
            s[f"{p}.conv1.weight"] = (ci, ci, 3, 3)
            s[f"{p}.conv1.bias"] = (ci,)
            s[f"{p}.conv2.weight"] = (co, ci, 3, 3)
            s[f"{p}.norm.weight"] = (co,)
            s[f"{p}.norm.bias"] = (co,)
        else:
            inner = int(ci // hd) * hd
            for nme in ("to_q", "to_k", "to_v"):
                s[f"{p}.attn.{nme}.weight"] = (inner, ci)
            s[f"{p}.attn.to_qkv_multiscale.0.proj_in.weight"] = (3 * inner, 1, 5, 5)
            s[f"{p}.attn.to_qkv_multiscale.0.proj_out.weight"] = (3 * inner, hd, 1, 1)
            s[f"{p}.attn.to_out.weight"] = (ci, 2 * inner)
            s[f"{p}.attn.norm_out.weight"] = (ci,)
            s[f"{p}.attn.norm_out.bias"] = (ci,)
            s[f"{p}.conv_out.conv_inverted.weight"] = (8 * ci, ci, 1, 1)
            s[f"{p}.conv_out.conv_inverted.bias"] = (8 * ci,)
            s[f"{p}.conv_out.conv_depth.weight"] = (8 * ci, 1, 3, 3)
            s[f"{p}.conv_out.conv_depth.bias"] = (8 * ci,)
            s[f"{p}.conv_out.conv_point.weight"] = (co, 4 * ci, 1, 1)
            s[f"{p}.conv_out.norm.weight"] = (co,)
            s[f"{p}.conv_out.norm.bias"] = (co,)
    s["decoder.norm_out.weight"] = (chans[0],)
    s["decoder.norm_out.bias"] = (chans[0],)
    s["decoder.conv_out.weight"] = (cfg["out_channels"], chans[0], 3, 3)
    s["decoder.conv_out.bias"] = (cfg["out_channels"],)
    return s


def dcae_encoder_layout(cfg: dict) -> List[Tuple[str, str, int, int]]:
    """Ordered encoder.down_blocks: (kind, key-prefix, C_in, C_out).  models/DCAE.py:582-606."""
    chans = cfg["encoder_block_out_channels"]
    layers = cfg["encoder_layers_per_block"]
    types = cfg["encoder_block_types"]
    out: List[Tuple[str, str, int, int]] = []
    j = 0
    n = len(chans)
    for i in range(n):
        for _ in range(layers[i]):
            kind = "res" if types[i] == "ResBlock" else "evit"
            out.append((kind, f"encoder.down_blocks.{j}", chans[i], chans[i]))
            j += 1
        if i < n - 1 and layers[i] > 0:
            out.append(("down", f"encoder.down_blocks.{j}", chans[i], chans[i + 1]))
            j += 1
    return out


def _block_param_shapes(s: Dict[str, Tuple[int, ...]], kind: str, p: str, ci: int, co: int, hd: int) -> None:
    if kind == "res":
        s[f"{p}.conv1.weight"] = (ci, ci, 3, 3)
        s[f"{p}.conv1.bias"] = (ci,)
        s[f"{p}.conv2.weight"] = (co, ci, 3, 3)
        s[f"{p}.norm.weight"] = (co,)
        s[f"{p}.norm.bias"] = (co,)
    else:
        inner = int(ci // hd) * hd
        for nme in ("to_q", "to_k", "to_v"):
            s[f"{p}.attn.{nme}.weight"] = (inner, ci)
        s[f"{p}.attn.to_qkv_multiscale.0.proj_in.weight"] = (3 * inner, 1, 5, 5)
        s[f"{p}.attn.to_qkv_multiscale.0.proj_out.weight"] = (3 * inner, hd, 1, 1)
        s[f"{p}.attn.to_out.weight"] = (ci, 2 * inner)
        s[f"{p}.attn.norm_out.weight"] = (ci,)
        s[f"{p}.attn.norm_out.bias"] = (ci,)
        s[f"{p}.conv_out.conv_inverted.weight"] = (8 * ci, ci, 1, 1)
        s[f"{p}.conv_out.conv_inverted.bias"] = (8 * ci,)
        s[f"{p}.conv_out.conv_depth.weight"] = (8 * ci, 1, 3, 3)
        s[f"{p}.conv_out.conv_depth.bias"] = (8 * ci,)
        s[f"{p}.conv_out.conv_point.weight"] = (co, 4 * ci, 1, 1)
        s[f"{p}.conv_out.norm.weight"] = (co,)
        s[f"{p}.conv_out.norm.bias"] = (co,)


def dcae_encoder_param_shapes(cfg: dict) -> Dict[str, Tuple[int, ...]]:
    """State-dict keys/shapes of AutoencoderDC.encoder (models/DCAE.py:539-615), layers_per_block[0] > 0 variant."""
    s: Dict[str, Tuple[int, ...]] = {}
    chans = cfg["encoder_block_out_channels"]
    assert cfg["encoder_layers_per_block"][0] > 0, "conv_in as DCDownBlock2d (layers_per_block[0] == 0) is not covered"
    s["encoder.conv_in.weight"] = (chans[0], cfg["in_channels"], 3, 3)
    s["encoder.conv_in.bias"] = (chans[0],)
    for kind, p, ci, co in dcae_encoder_layout(cfg):
        if kind == "down":
            s[f"{p}.conv.weight"] = (co // 4, ci, 3, 3)
            s[f"{p}.conv.bias"] = (co // 4,)
        else:
            _block_param_shapes(s, kind, p, ci, co, cfg["attention_head_dim"])
    s["encoder.conv_out.weight"] = (cfg["latent_channels"], chans[-1], 3, 3)
    s["encoder.conv_out.bias"] = (cfg["latent_channels"],)
    return s


def make_state_dict(shapes: Dict[str, Tuple[int, ...]], salt: int = 0) -> SD:
    return {k: det_tensor(k, shp, salt) for k, shp in shapes.items()}


# --------------------------------------------------------------------------------------------------------------
# small primitives (diffusers restatements)
# --------------------------------------------------------------------------------------------------------------


def _lin(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def _ln(x: torch.Tensor, eps: float, w=None, b=None) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def _rms(x: torch.Tensor, eps: float, w=None, b=None) -> torch.Tensor:
    var = x.float().pow(2).mean(-1, keepdim=True)
    x = x * torch.rsqrt(var + eps)
    if w is not None:
        x = x * w
    if b is not None:
        x = x + b
    return x


def timestep_sincos(t: torch.Tensor, dim: int = 256) -> torch.Tensor:
    """diffusers Timesteps(256, flip_sin_to_cos=True, downscale_freq_shift=0): [cos | sin]."""
    half = dim // 2
    f = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half)
    arg = t[:, None].float() * f[None]
    return torch.cat([torch.cos(arg), torch.sin(arg)], dim=-1)


def combined_time_text(sd: SD, pre: str, t: torch.Tensor, pooled: torch.Tensor) -> torch.Tensor:
    """CombinedTimestepTextProjEmbeddings.forward (used LaDCast_3D_model.py:362,673)."""
    te = _lin(sd, pre + ".timestep_embedder.linear_2", F.silu(_lin(sd, pre + ".timestep_embedder.linear_1",
                                                                   timestep_sincos(t))))
    pe = _lin(sd, pre + ".text_embedder.linear_2", F.silu(_lin(sd, pre + ".text_embedder.linear_1", pooled)))
    return te + pe


def rope_1d(dim: int, pos: torch.Tensor, theta: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """get_1d_rotary_pos_embed(use_real=True): each frequency repeated twice (interleaved)."""
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float32)[: dim // 2] / dim))
    ang = torch.outer(pos.float(), freqs)
    return ang.cos().repeat_interleave(2, dim=1), ang.sin().repeat_interleave(2, dim=1)


def rope_tables(cfg: dict, T_in: int, T_out: int, H: int, W: int):
    """RoPE cos/sin tables [T*H*W, head_dim] for pred and cond tokens.
    LaDCast_3D_model.py:885-938 + embeddings.py:274-327 (meshgrid 'ij' over t, lat, lon; per-axis tables
    concatenated along the feature dim)."""
    start = cfg["rope_spatial_grid_start_pos"]
    end = cfg["rope_spatial_grid_end_pos"]
    if cfg.get("spatial_deg2rad", False):
        start = [float(np.deg2rad(v)) for v in start]
        end = [float(np.deg2rad(v)) for v in end]
    lat = torch.linspace(start[0], end[0], steps=H, dtype=torch.float32)
    lon = torch.linspace(start[1], end[1], steps=W, dtype=torch.float32)
    theta = cfg["rope_theta"]

    def table(tc, dims):
        g = torch.stack(torch.meshgrid(tc, lat, lon, indexing="ij"), dim=0)
        cs = [rope_1d(dims[i], g[i].reshape(-1), theta) for i in range(3)]
        return torch.cat([c for c, _ in cs], dim=1), torch.cat([s for _, s in cs], dim=1)

    t_cond = torch.arange(-T_in + 1, 1, dtype=torch.float32)
    t_pred = torch.arange(1, T_out + 1, dtype=torch.float32)
    return table(t_pred, cfg["rope_axes_dim"]), table(t_cond, cfg["conditioning_tensor_rope_axes_dim"])


def apply_rope(x: torch.Tensor, cs) -> torch.Tensor:
    """apply_rotary_emb, interleaved pairs: out = x*cos + rot(x)*sin, rot(x0,x1) = (-x1, x0).  x: [B,H,S,D]."""
    cos, sin = cs
    xr, xi = x.reshape(*x.shape[:-1], -1, 2).unbind(-1)
    rot = torch.stack([-xi, xr], dim=-1).flatten(3)
    return (x.float() * cos[None, None] + rot.float() * sin[None, None]).to(x.dtype)


def year_progress(ts: int) -> float:
    """embeddings.py:428-447: YYYYMMDDHH -> fraction of the (leap-aware) year elapsed."""
    s = str(int(ts))
    dt = datetime(int(s[0:4]), int(s[4:6]), int(s[6:8]), int(s[8:10]))
    y0, y1 = datetime(dt.year, 1, 1), datetime(dt.year + 1, 1, 1)
    return (dt - y0).total_seconds() / (y1 - y0).total_seconds()


def year_sincos(ts: torch.Tensor, dim: int = 256) -> torch.Tensor:
    """get_year_sincos_embedding (embeddings.py:467-520)."""
    half = dim // 2
    p = torch.tensor([year_progress(int(v)) for v in ts.tolist()], dtype=torch.float32)
    k = torch.arange(1, half + 1).float()
    mag = torch.exp(-math.log(10000) * torch.arange(0, half).float() / half)
    arg = (2 * math.pi * p.reshape(-1, 1)) * k.reshape(1, -1)
    return torch.cat([torch.sin(arg) * mag[None], torch.cos(arg) * mag[None]], dim=1)


def _heads(x: torch.Tensor, nh: int) -> torch.Tensor:
    return x.unflatten(2, (nh, -1)).transpose(1, 2)


def _merge(x: torch.Tensor) -> torch.Tensor:
    return x.transpose(1, 2).flatten(2, 3)


# --------------------------------------------------------------------------------------------------------------
# denoiser (models/LaDCast_3D_model.py)
# --------------------------------------------------------------------------------------------------------------


def patch_embed(sd: SD, pre: str, x: torch.Tensor) -> torch.Tensor:
    """HunyuanVideoPatchEmbed with patch (1,1,1) (embeddings.py:38-59): token n = t*H*W + h*W + w."""
    w = sd[pre + ".proj.weight"].flatten(1)
    tok = x.flatten(2).transpose(1, 2)
    return F.linear(tok, w, sd[pre + ".proj.bias"])


def refiner_block(sd: SD, p: str, x, temb, rope_c, nh: int):
    """LaDCastIndividualTokenRefinerBlock.forward (LaDCast_3D_model.py:280-302); attention is pre_only."""
    n1 = _ln(x, 1e-7, sd[p + ".norm1.weight"], sd[p + ".norm1.bias"])
    q = _heads(_lin(sd, p + ".attn.to_q", n1), nh)
    k = _heads(_lin(sd, p + ".attn.to_k", n1), nh)
    v = _heads(_lin(sd, p + ".attn.to_v", n1), nh)
    q = apply_rope(_rms(q, 1e-7, sd[p + ".attn.norm_q.weight"]), rope_c)
    k = apply_rope(_rms(k, 1e-7, sd[p + ".attn.norm_k.weight"]), rope_c)
    a = _merge(F.scaled_dot_product_attention(q, k, v))
    g = _lin(sd, p + ".norm_out.linear", F.silu(temb))
    g_msa, g_mlp = g.chunk(2, dim=1)
    x = x + a * g_msa[:, None]
    n2 = _ln(x, 1e-7, sd[p + ".norm2.weight"], sd[p + ".norm2.bias"])
    ff = _lin(sd, p + ".ff.net.2", F.silu(_lin(sd, p + ".ff.net.0.proj", n2)))
    return x + ff * g_mlp[:, None]


def dual_block(sd: SD, p: str, h, e, temb, rope_p, nh: int):
    """LaDCastTransformerBlock.forward (LaDCast_3D_model.py:514-566) + processor (:78-221) in the dual-stream
    branch: RoPE on pred tokens only; cond q/k get RMSNorm only."""
    m = _lin(sd, p + ".norm1.linear", F.silu(temb)).chunk(6, dim=1)
    mc = _lin(sd, p + ".norm1_context.linear", F.silu(temb)).chunk(6, dim=1)
    nh_ = _ln(h, 1e-6) * (1 + m[1][:, None]) + m[0][:, None]
    ne_ = _ln(e, 1e-6) * (1 + mc[1][:, None]) + mc[0][:, None]
    q = apply_rope(_rms(_heads(_lin(sd, p + ".attn.to_q", nh_), nh), 1e-7, sd[p + ".attn.norm_q.weight"]), rope_p)
    k = apply_rope(_rms(_heads(_lin(sd, p + ".attn.to_k", nh_), nh), 1e-7, sd[p + ".attn.norm_k.weight"]), rope_p)
    v = _heads(_lin(sd, p + ".attn.to_v", nh_), nh)
    eq = _rms(_heads(_lin(sd, p + ".attn.add_q_proj", ne_), nh), 1e-7, sd[p + ".attn.norm_added_q.weight"])
    ek = _rms(_heads(_lin(sd, p + ".attn.add_k_proj", ne_), nh), 1e-7, sd[p + ".attn.norm_added_k.weight"])
    ev = _heads(_lin(sd, p + ".attn.add_v_proj", ne_), nh)
    a = _merge(F.scaled_dot_product_attention(torch.cat([q, eq], 2), torch.cat([k, ek], 2), torch.cat([v, ev], 2)))
    Np = h.shape[1]
    ah = _lin(sd, p + ".attn.to_out.0", a[:, :Np])
    ae = _lin(sd, p + ".attn.to_add_out", a[:, Np:])
    h = h + ah * m[2][:, None]
    e = e + ae * mc[2][:, None]
    n2h = _ln(h, 1e-7) * (1 + m[4][:, None]) + m[3][:, None]
    n2e = _ln(e, 1e-7) * (1 + mc[4][:, None]) + mc[3][:, None]
    fh = _lin(sd, p + ".ff.net.2", F.gelu(_lin(sd, p + ".ff.net.0.proj", n2h), approximate="tanh"))
    fe = _lin(sd, p + ".ff_context.net.2", F.gelu(_lin(sd, p + ".ff_context.net.0.proj", n2e), approximate="tanh"))
    return h + m[5][:, None] * fh, e + mc[5][:, None] * fe


def single_block(sd: SD, p: str, h, e, temb, rope_p, rope_c, nh: int):
    """LaDCastSingleTransformerBlock.forward (LaDCast_3D_model.py:426-468): streams concatenated, parallel
    attention + MLP, pred and cond tokens rotated with their own tables (:115-141), no attention out-proj."""
    Np = h.shape[1]
    x = torch.cat([h, e], dim=1)
    shift, scale, gate = _lin(sd, p + ".norm.linear", F.silu(temb)).chunk(3, dim=1)
    n = _ln(x, 1e-6) * (1 + scale[:, None]) + shift[:, None]
    mlp = F.gelu(_lin(sd, p + ".proj_mlp", n), approximate="tanh")
    q = _rms(_heads(_lin(sd, p + ".attn.to_q", n), nh), 1e-7, sd[p + ".attn.norm_q.weight"])
    k = _rms(_heads(_lin(sd, p + ".attn.to_k", n), nh), 1e-7, sd[p + ".attn.norm_k.weight"])
    v = _heads(_lin(sd, p + ".attn.to_v", n), nh)
    q = torch.cat([apply_rope(q[:, :, :Np], rope_p), apply_rope(q[:, :, Np:], rope_c)], dim=2)
    k = torch.cat([apply_rope(k[:, :, :Np], rope_p), apply_rope(k[:, :, Np:], rope_c)], dim=2)
    a = _merge(F.scaled_dot_product_attention(q, k, v))
    y = gate[:, None] * _lin(sd, p + ".proj_out", torch.cat([a, mlp], dim=2)) + x
    return y[:, :Np], y[:, Np:]


def denoiser_forward(sd: SD, cfg: dict, x: torch.Tensor, timestep: torch.Tensor, cond: torch.Tensor,
                     time_elapsed: Optional[torch.Tensor] = None, taps: Optional[dict] = None) -> torch.Tensor:
    """LaDCastTransformer3DModel.forward (LaDCast_3D_model.py:833-1071), patch_size = patch_size_t = 1.
    x [B,C,T_out,H,W]; timestep (B,) or (1,) float (= 0.25 ln sigma); cond [B,C,T_in,H,W]; time_elapsed int64
    (1,) or (B,) YYYYMMDDHH.  Returns [B,C_out,T_out,H,W]."""
    B, _, T_out, H, W = x.shape
    T_in = cond.shape[2]
    nh = cfg["num_attention_heads"]
    if timestep.numel() == 1 and B > 1:
        timestep = timestep.reshape(-1).expand(B)
    rope_p, rope_c = rope_tables(cfg, T_in, T_out, H, W)
    h = patch_embed(sd, "x_embedder", x)
    e = patch_embed(sd, "context_embedder", cond)
    # context refiner (:375-390)
    r_temb = combined_time_text(sd, "context_refiner.time_text_embed", timestep, e.mean(dim=1))
    e = _lin(sd, "context_refiner.proj_in", e)
    for i in range(cfg["num_refiner_layers"]):
        e = refiner_block(sd, f"context_refiner.token_refiner.refiner_blocks.{i}", e, r_temb, rope_c, nh)
    if taps is not None:
        taps["refined_cond"] = e.clone()
    temb = combined_time_text(sd, "time_text_embed", timestep, e.mean(dim=1))
    if time_elapsed is not None and cfg.get("incl_time_elapsed", False):
        ye = year_sincos(time_elapsed.reshape(-1))
        ye = _lin(sd, "time_elapsed_embed.linear_2", F.silu(_lin(sd, "time_elapsed_embed.linear_1", ye)))
        sc, sh = ye.chunk(2, dim=-1)
        temb = temb * (1 + sc) + sh
    if taps is not None:
        taps["temb"] = temb.clone()
    for i in range(cfg["num_layers"]):
        h, e = dual_block(sd, f"transformer_blocks.{i}", h, e, temb, rope_p, nh)
        if taps is not None:
            taps[f"dual{i}.h"] = h.clone()
            taps[f"dual{i}.e"] = e.clone()
    for i in range(cfg["num_single_layers"]):
        h, e = single_block(sd, f"single_transformer_blocks.{i}", h, e, temb, rope_p, rope_c, nh)
        if taps is not None:
            taps[f"single{i}.h"] = h.clone()
    scale, shift = _lin(sd, "norm_out.linear", F.silu(temb)).chunk(2, dim=1)
    h = _ln(h, 1e-7) * (1 + scale)[:, None] + shift[:, None]
    out = _lin(sd, "proj_out", h)  # [B, T*H*W, C_out]
    return out.reshape(B, T_out, H, W, -1).permute(0, 4, 1, 2, 3).contiguous()


# --------------------------------------------------------------------------------------------------------------
# sphere convolution + DC-AE decoder (models/sphere_conv.py, models/DCAE.py)
# --------------------------------------------------------------------------------------------------------------


def sphere_pad(x: torch.Tensor, ph: int, pw: int) -> torch.Tensor:
    """SphereConv2d.sphere_pad (sphere_conv.py:62-91): pole rows = first/last `ph` rows rolled by W/2 and
    flipped vertically; longitude circular."""
    half = x.shape[3] // 2
    top = torch.flip(torch.roll(x[:, :, :ph], half, dims=3), dims=[2])
    bot = torch.flip(torch.roll(x[:, :, -ph:], half, dims=3), dims=[2])
    x = torch.cat([top, x, bot], dim=2)
    return F.pad(x, (pw, pw, 0, 0), mode="circular")


def sphere_conv(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], groups: int = 1) -> torch.Tensor:
    """SphereConv2d.forward (sphere_conv.py:138-192): output row 0 / H-1 use a kernel whose first / last `pad`
    rows are mirrored left-right (:93-129); all other rows use the plain kernel on the padded input."""
    k = w.shape[-1]
    p = k // 2
    xp = sphere_pad(x, p, p)
    w_top = w.clone()
    w_top[:, :, :p, :] = torch.flip(w[:, :, :p, :], dims=[3])
    w_bot = w.clone()
    w_bot[:, :, -p:, :] = torch.flip(w[:, :, -p:, :], dims=[3])
    top = F.conv2d(xp[:, :, :k], w_top, b, 1, 0, 1, groups)
    mid = F.conv2d(xp[:, :, 1:-1], w, b, 1, 0, 1, groups)
    bot = F.conv2d(xp[:, :, -k:], w_bot, b, 1, 0, 1, groups)
    return torch.cat([top, mid, bot], dim=2)


def _rms_c(x: torch.Tensor, eps: float, w, b) -> torch.Tensor:
    return _rms(x.movedim(1, -1), eps, w, b).movedim(-1, 1)


def res_block(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """ResBlock.forward (DCAE.py:356-377), norm = RMSNorm over channels eps 1e-5 (get_normalization default)."""
    y = F.silu(sphere_conv(x, sd[p + ".conv1.weight"], sd[p + ".conv1.bias"]))
    y = sphere_conv(y, sd[p + ".conv2.weight"], None)
    return _rms_c(y, 1e-5, sd[p + ".norm.weight"], sd[p + ".norm.bias"]) + x


def evit_block(sd: SD, p: str, x: torch.Tensor, hd: int) -> torch.Tensor:
    """EfficientViTBlock = SanaMultiscaleLinearAttention (processor DCAE.py:205-267, linear attention
    :155-175) followed by GLUMBConv (:304-324)."""
    n, C, H, W = x.shape
    a = p + ".attn"
    t = x.movedim(1, -1)
    qkv = torch.cat([F.linear(t, sd[a + ".to_q.weight"]), F.linear(t, sd[a + ".to_k.weight"]),
                     F.linear(t, sd[a + ".to_v.weight"])], dim=3).movedim(-1, 1)
    inner3 = qkv.shape[1]
    ms = sphere_conv(qkv, sd[a + ".to_qkv_multiscale.0.proj_in.weight"], None, groups=inner3)
    ms = F.conv2d(ms, sd[a + ".to_qkv_multiscale.0.proj_out.weight"], None, 1, 0, 1, inner3 // hd)
    hs = torch.cat([qkv, ms], dim=1).float().reshape(n, -1, 3 * hd, H * W)
    q, k, v = hs.chunk(3, dim=2)
    q, k = F.relu(q), F.relu(k)
    if H * W > hd:
        v1 = F.pad(v, (0, 0, 0, 1), mode="constant", value=1)
        o = torch.matmul(torch.matmul(v1, k.transpose(-1, -2)), q)
        o = o[:, :, :-1] / (o[:, :, -1:] + 1e-15)
    else:
        sc = torch.matmul(k.transpose(-1, -2), q)
        sc = sc / (sc.sum(dim=2, keepdim=True) + 1e-15)
        o = torch.matmul(v, sc)
    o = o.reshape(n, -1, H, W)
    o = F.linear(o.movedim(1, -1), sd[a + ".to_out.weight"]).movedim(-1, 1)
    x = _rms_c(o, 1e-5, sd[a + ".norm_out.weight"], sd[a + ".norm_out.bias"]) + x
    c = p + ".conv_out"
    y = F.silu(F.conv2d(x, sd[c + ".conv_inverted.weight"], sd[c + ".conv_inverted.bias"]))
    y = sphere_conv(y, sd[c + ".conv_depth.weight"], sd[c + ".conv_depth.bias"], groups=y.shape[1])
    y, gate = y.chunk(2, dim=1)
    y = F.conv2d(y * F.silu(gate), sd[c + ".conv_point.weight"], None)
    return _rms_c(y, 1e-7, sd[c + ".norm.weight"], sd[c + ".norm.bias"]) + x


def up_block(sd: SD, p: str, x: torch.Tensor, c_out: int) -> torch.Tensor:
    """DCUpBlock2d.forward, pixel-shuffle variant with shortcut (DCAE.py:519-536)."""
    y = F.pixel_shuffle(sphere_conv(x, sd[p + ".conv.weight"], sd[p + ".conv.bias"]), 2)
    rep = c_out * 4 // x.shape[1]
    return y + F.pixel_shuffle(x.repeat_interleave(rep, dim=1), 2)


def down_block(sd: SD, p: str, x: torch.Tensor, c_out: int) -> torch.Tensor:
    """DCDownBlock2d.forward, pixel-unshuffle variant with the channel-averaging shortcut (DCAE.py:476-490)."""
    y = F.pixel_unshuffle(sphere_conv(x, sd[p + ".conv.weight"], sd[p + ".conv.bias"]), 2)
    g = x.shape[1] * 4 // c_out
    return y + F.pixel_unshuffle(x, 2).unflatten(1, (-1, g)).mean(dim=2)


def dcae_encode(sd: SD, cfg: dict, x: torch.Tensor, static: Optional[torch.Tensor] = None,
                taps: Optional[dict] = None) -> torch.Tensor:
    """AutoencoderDC.encode -> Encoder.forward (DCAE.py:964-1000, 617-631).  x [n,84(+5),H,W] -> [n,84,H/8,W/8]."""
    if static is not None:
        x = torch.cat((x, static), dim=1)
    h = sphere_conv(x, sd["encoder.conv_in.weight"], sd["encoder.conv_in.bias"])
    if taps is not None:
        taps["conv_in"] = h.clone()
    for kind, p, ci, co in dcae_encoder_layout(cfg):
        if kind == "down":
            h = down_block(sd, p, h, co)
        elif kind == "res":
            h = res_block(sd, p, h)
        else:
            h = evit_block(sd, p, h, cfg["attention_head_dim"])
        if taps is not None:
            taps[p] = h.clone()
    g = cfg["encoder_block_out_channels"][-1] // cfg["latent_channels"]
    return sphere_conv(h, sd["encoder.conv_out.weight"], sd["encoder.conv_out.bias"]) + h.unflatten(1, (-1, g)).mean(dim=2)


def dcae_decode(sd: SD, cfg: dict, z: torch.Tensor, return_static: bool = False,
                taps: Optional[dict] = None) -> torch.Tensor:
    """AutoencoderDC.decode -> Decoder.forward (DCAE.py:1018-1056, 717-732).  z [n,84,h,w] -> [n,84,8h,8w]."""
    chans = cfg["decoder_block_out_channels"]
    rep = chans[-1] // cfg["latent_channels"]
    x = sphere_conv(z, sd["decoder.conv_in.weight"], sd["decoder.conv_in.bias"]) + z.repeat_interleave(rep, dim=1)
    if taps is not None:
        taps["conv_in"] = x.clone()
    for kind, p, ci, co in dcae_decoder_layout(cfg):
        if kind == "up":
            x = up_block(sd, p, x, co)
        elif kind == "res":
            x = res_block(sd, p, x)
        else:
            x = evit_block(sd, p, x, cfg["attention_head_dim"])
        if taps is not None:
            taps[p] = x.clone()
    x = F.relu(_rms_c(x, 1e-7, sd["decoder.norm_out.weight"], sd["decoder.norm_out.bias"]))
    x = sphere_conv(x, sd["decoder.conv_out.weight"], sd["decoder.conv_out.bias"])
    if not return_static and cfg.get("static_channels"):
        x = x[:, : -cfg["static_channels"]]
    return x


def decode_latent_ens(sd: SD, cfg: dict, latents: torch.Tensor, mean=None, std=None) -> torch.Tensor:
    """pipelines/utils.py:52-80: [B,C,T,h,w] -> decode each (b,t) frame -> [B,84,T,H,W] (*std + mean)."""
    B, C, T, h, w = latents.shape
    y = dcae_decode(sd, cfg, latents.permute(0, 2, 1, 3, 4).reshape(B * T, C, h, w))
    y = y.reshape(B, T, y.shape[1], y.shape[2], y.shape[3]).permute(0, 2, 1, 3, 4)
    if mean is not None:
        y = y * std[None, :, None, None, None] + mean[None, :, None, None, None]
    return y


# --------------------------------------------------------------------------------------------------------------
# scheduler + samplers (diffusers EDMDPMSolverMultistepScheduler; pipelines/pipeline_AR.py; edm_sampler.py)
# --------------------------------------------------------------------------------------------------------------

SIGMA_MIN, SIGMA_MAX, SIGMA_DATA, RHO = 0.002, 80.0, 0.5, 7.0


def karras_sigmas(n: int) -> torch.Tensor:
    """set_timesteps: N Karras sigmas (f32) followed by a final 0."""
    ramp = torch.linspace(0, 1, n)
    lo, hi = SIGMA_MIN ** (1 / RHO), SIGMA_MAX ** (1 / RHO)
    s = ((hi + ramp * (lo - hi)) ** RHO).to(torch.float32)
    return torch.cat([s, torch.zeros(1)])


def precondition(sig):
    """c_in, c_skip, c_out, c_noise for sigma (0-dim tensor): precondition_inputs/outputs/noise."""
    sd2 = SIGMA_DATA**2
    c_in = 1 / ((sig**2 + sd2) ** 0.5)
    c_skip = sd2 / (sig**2 + sd2)
    c_out = sig * SIGMA_DATA / (sig**2 + sd2) ** 0.5
    return c_in, c_skip, c_out, 0.25 * torch.log(sig)


def dpmpp2m_sample(net, noise: torch.Tensor, n_steps: int, trace: Optional[list] = None) -> torch.Tensor:
    """AutoRegressive2DPipeline.__call__ loop (pipeline_AR.py:84-102) with scheduler.scale_model_input / .step
    inlined.  With solver_order=2 and final_sigmas_type="zero" the update is first-order at i=0 and at i=N-1 and
    the 2M midpoint rule otherwise (`lower_order_second` only matters for order 3).  net(x_in, c_noise[B]) -> F.
    The initial noise is NOT scaled by sigma_0 (reference quirk, pipeline_AR.py:77-82)."""
    sig = karras_sigmas(n_steps)
    x = noise
    prev_x0 = None
    B = x.shape[0]
    for i in range(n_steps):
        s, s_next = sig[i], sig[i + 1]
        c_in, c_skip, c_out, c_noise = precondition(s)
        f = net(x * c_in, c_noise.expand(B))
        x0 = c_skip * x + c_out * f
        ratio = s_next / s
        h = -torch.log(s_next) + torch.log(s)
        em1 = torch.exp(-h) - 1.0
        if i == 0 or i == n_steps - 1:
            x = ratio * x - em1 * x0
        else:
            h0 = -torch.log(s) + torch.log(sig[i - 1])
            r0 = h0 / h
            d1 = (1.0 / r0) * (x0 - prev_x0)
            x = ratio * x - em1 * x0 - 0.5 * em1 * d1
        prev_x0 = x0
        if trace is not None:
            trace.append(x.clone())
    return x


def heun_sample(net, noise: torch.Tensor, n_steps: int, deterministic: bool = True, S_churn=0.0, S_min=0.0,
                S_max=float("inf"), S_noise=0.0, randn_like=torch.randn_like) -> torch.Tensor:
    """edm_AR_sampler (edm_sampler.py:44-120): float64 state, 2N-1 net calls with a (1,) c_noise broadcast over the
    batch; deterministic=False adds the stochastic churn of :67-76 (t_hat, gamma in float32)."""
    t_steps = karras_sigmas(n_steps)
    x_next = noise.to(torch.float64) * t_steps[0]

    def den(xx, t):
        c_in, c_skip, c_out, c_noise = precondition(t)
        f = net((xx * c_in).to(torch.float32), c_noise.reshape(-1).to(torch.float32)).to(torch.float64)
        return c_skip * xx + c_out * f

    for i in range(n_steps):
        t_cur, t_nxt = t_steps[i], t_steps[i + 1]
        x_hat, t_hat = x_next, t_cur
        if not deterministic:
            gamma = min(S_churn / n_steps, 2.0**0.5 - 1) if S_min <= float(t_cur) <= S_max else 0
            t_hat = t_cur + gamma * t_cur
            x_hat = x_next + (t_hat**2 - t_cur**2).sqrt() * S_noise * randn_like(x_next)
        d_cur = (x_hat - den(x_hat, t_hat)) / t_hat
        x_next = x_hat + (t_nxt - t_hat) * d_cur
        if i < n_steps - 1:
            d_prime = (x_next - den(x_next, t_nxt)) / t_nxt
            x_next = x_hat + (t_nxt - t_hat) * (0.5 * d_cur + 0.5 * d_prime)
    return x_next.float()


def member_noise(members: Sequence[int], shape_tail: Sequence[int], dtype=torch.float32) -> torch.Tensor:
    """ensemble_AR_sampler seeds (pipelines/utils.py:703-706) + randn_tensor list branch: member m draws
    torch.randn((1,*tail)) from torch.Generator('cpu').manual_seed(m) — the SAME seed at every AR step."""
    outs = []
    for m in members:
        g = torch.Generator("cpu").manual_seed(int(m) % (1 << 32))
        outs.append(torch.randn((1, *shape_tail), generator=g, dtype=dtype))
    return torch.cat(outs, dim=0)


def ensemble_sample(sd: SD, cfg: dict, members: Sequence[int], T_out: int, n_steps: int, known: torch.Tensor,
                    timestamp: int, sampler: str = "pipeline", sampler_kwargs: Optional[dict] = None) -> torch.Tensor:
    """ensemble_AR_sampler (pipelines/utils.py:665-742) for an explicit list of global member indices."""
    B = len(members)
    kn = known.expand(B, *known.shape[1:]) if known.shape[0] == 1 else known
    ts = torch.tensor([timestamp], dtype=torch.int64)
    noise = member_noise(members, (cfg["out_channels"], T_out, *known.shape[-2:]))
    net = lambda xin, cn: denoiser_forward(sd, cfg, xin, cn, kn, ts)  # noqa: E731
    if sampler == "pipeline":
        return dpmpp2m_sample(net, noise, n_steps)
    if sampler == "edm":
        return heun_sample(net, noise, n_steps, **(sampler_kwargs or {}))
    raise ValueError(sampler)


def normalize_latent(x, mean, std, target_std=0.5):
    """normalize_transform_3D (dataloader/utils.py:223-230), channel dim = -4."""
    return (x - mean[:, None, None, None]) / std[:, None, None, None] * target_std


def denormalize_latent(x, mean, std, target_std=0.5):
    """inverse_normalize_transform_3D (dataloader/utils.py:233-240)."""
    return (x / target_std) * std[:, None, None, None] + mean[:, None, None, None]


def advance_timestamp(ts: int, hours: int) -> int:
    """pipelines/utils.py:538-541: init + step*6h*T_out, formatted back to YYYYMMDDHH."""
    import datetime as _dt

    s = str(int(ts))
    d = datetime(int(s[0:4]), int(s[4:6]), int(s[6:8]), int(s[8:10])) + _dt.timedelta(hours=hours)
    return int(d.strftime("%Y%m%d%H"))


def rollout(den_sd: SD, den_cfg: dict, ae_sd: Optional[SD], ae_cfg: Optional[dict], known: torch.Tensor,
            members: Sequence[int], init_ts: int, total_steps: int, T_out: int, n_steps: int, lat_mean, lat_std,
            field_mean=None, field_std=None, sampler: str = "pipeline", decode: bool = True):
    """roll_out_serial loop body (pipelines/utils.py:533-585) from already-normalised known latents
    [1,C,T_in,h,w].  Returns (latents [B,C,total,h,w] de-normalised, fields [B,84,total,H,W] or None)."""
    T_in = known.shape[2]
    reps = math.ceil(total_steps / T_out)
    lat_out, fld_out = [], []
    for step in range(reps):
        cur = min(1 + (step + 1) * T_out, total_steps + 1)
        sel = cur - (1 + step * T_out)
        ts = advance_timestamp(init_ts, step * 6 * T_out)
        s = ensemble_sample(den_sd, den_cfg, members, T_out, n_steps, known, ts, sampler)
        known = s[:, :, -T_in:].clone()
        phys = denormalize_latent(s, lat_mean, lat_std)
        lat_out.append(phys[:, :, :sel])
        if decode and ae_sd is not None:
            fld_out.append(decode_latent_ens(ae_sd, ae_cfg, phys[:, :, :sel], field_mean, field_std))
    return torch.cat(lat_out, dim=2), (torch.cat(fld_out, dim=2) if fld_out else None)


# --------------------------------------------------------------------------------------------------------------
# metrics (evaluate/utils.py:40-118; assembly evaluate/evaluate_ens_gpu.py:339-415)
# --------------------------------------------------------------------------------------------------------------

SST_CHANNEL = 82


def lat_weights(n_lat: int = 120) -> np.ndarray:
    """get_normalized_lat_weights_based_on_cos on linspace(-88.5, 90, 120) — numpy float64
    (evaluate_ens_gpu.py:166-168)."""
    w = np.cos(np.deg2rad(np.linspace(-88.5, 90.0, n_lat)))
    return w / w.mean()


def crps_spread_pointwise(fc: torch.Tensor) -> torch.Tensor:
    """pointwise_crps_spread (evaluate/utils.py:63-101), ensemble dim 0."""
    M = fc.shape[0]
    if M < 2:
        return torch.zeros_like(fc[0])
    srt, _ = torch.sort(fc, dim=0)
    wts = (2 * torch.arange(1, M + 1, dtype=fc.dtype) - M - 1).view(-1, *([1] * (fc.ndim - 1)))
    return 2 * (srt * wts).sum(dim=0) / (M * (M - 1))


def ensemble_metrics(fields: torch.Tensor, truth: torch.Tensor) -> Dict[str, torch.Tensor]:
    """Per-(channel, lead) tables of ens-mean MSE, CRPS skill, spread, total — evaluate_ens_gpu.py:339-415.
    fields [M,C,T,H,W] f32; truth [C,T,H,W] f32 (NaN allowed in the SST channel); weights float64 so products
    are float64; channel 82 reduced with nanmean, others with mean."""
    M, C, T, H, W = fields.shape
    w = torch.from_numpy(lat_weights(H)).view(1, -1, 1)
    out = {k: torch.zeros(C, T, dtype=torch.float64) for k in ("ens_mse", "crps_skill", "crps_spread", "crps")}

    def red(x):
        r = x.mean(dim=(1, 2))
        if C > SST_CHANNEL:
            r[SST_CHANNEL] = torch.nanmean(x[SST_CHANNEL : SST_CHANNEL + 1], dim=(1, 2))[0]
        return r

    for t in range(T):
        dec, ref = fields[:, :, t], truth[:, t]
        mean_t = dec.mean(dim=0)
        se = (mean_t - ref) ** 2 * w
        spread = crps_spread_pointwise(dec) * w
        skill = torch.abs(ref.unsqueeze(0) - dec).mean(dim=0) * w
        out["ens_mse"][:, t] = red(se)
        out["crps_spread"][:, t] = red(spread)
        out["crps_skill"][:, t] = red(skill)
        out["crps"][:, t] = red(skill - 0.5 * spread)
    return out


def get_acc(forecast: torch.Tensor, truth: torch.Tensor, climate: torch.Tensor, lat_weight=None) -> torch.Tensor:
    """get_acc (evaluate/utils.py:122-149): anomaly correlation over the last two dims with nanmean; lat_weight
    broadcastable [H, 1] (float64 in the scripts, so the products are float64)."""
    fa, ta = forecast - climate, truth - climate
    if lat_weight is None:
        num = torch.nanmean(fa * ta, dim=(-2, -1))
        den = torch.sqrt(torch.nanmean(fa**2, dim=(-2, -1)) * torch.nanmean(ta**2, dim=(-2, -1)))
    else:
        num = torch.nanmean(fa * ta * lat_weight, dim=(-2, -1))
        den = torch.sqrt(torch.nanmean(fa**2 * lat_weight, dim=(-2, -1)) * torch.nanmean(ta**2 * lat_weight, dim=(-2, -1)))
    return num / den
