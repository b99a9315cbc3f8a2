"""Drop-in for `ladcast.models.LaDCast_3D_model.LaDCastTransformer3DModel` (reference
models/LaDCast_3D_model.py:569-1071): same constructor keywords, `forward` signature, `.config`, checkpoint
layout and state-dict key names — but the forward pass runs entirely in the sm_100a CUDA library
(`lc_denoiser_*` in include/ladcast_b200.h).  There is no PyTorch/CPU fallback."""
import ctypes
import math
import os
from contextlib import contextmanager
from dataclasses import dataclass
from typing import Any, Dict, Optional, Tuple

import torch

from .. import _lib
from .embeddings import rope_tables, year_sincos_embedding
from .modeling import CheckpointMixin, capture_config


@dataclass
class Transformer2DModelOutput:
    sample: torch.Tensor

    def __getitem__(self, i):
        return (self.sample,)[i]


def _linear_init(shape, gen=None):
    fan_in = int(torch.tensor(shape[1:]).prod()) if len(shape) > 1 else shape[0]
    bound = 1.0 / math.sqrt(fan_in)
    return (torch.rand(shape, generator=gen) * 2 - 1) * bound


class LaDCastTransformer3DModel(CheckpointMixin):
    _class_name = "LaDCastTransformer3DModel"

    def __init__(self, in_channels: int = 16, out_channels: int = 16, num_attention_heads: int = 24,
                 attention_head_dim: int = 128, num_layers: int = 20, num_single_layers: int = 40,
                 num_refiner_layers: int = 2, mlp_ratio: float = 4.0, patch_size: int = 1, patch_size_t: int = 1,
                 qk_norm: str = "rms_norm", rope_theta: float = 256.0, rope_axes_dim: Tuple[int] = (16, 56, 56),
                 rope_spatial_grid_start_pos=0, rope_spatial_grid_end_pos=None, spatial_deg2rad: bool = False,
                 conditioning_tensor_in_channels: int = None, conditioning_tensor_intermediate_proj_dim: Optional[int] = None,
                 conditioning_tensor_rope_axes_dim: Tuple[int] = (16, 56, 56), incl_time_elapsed: bool = False,
                 nope: bool = False, scale_attn_by_lat: bool = False) -> None:
        kw = dict(locals())
        kw.pop("self")
        self.config = capture_config(type(self), kw)
        if nope or scale_attn_by_lat:
            raise NotImplementedError("nope / scale_attn_by_lat variants are not part of the V0.1.X checkpoints")
        if patch_size != 1 or patch_size_t != 1:
            raise NotImplementedError("only patch_size = patch_size_t = 1 (the shipped configs) is supported")
        if qk_norm != "rms_norm":
            raise NotImplementedError("qk_norm must be 'rms_norm'")
        assert sum(rope_axes_dim) == attention_head_dim, "sum(rope_axes_dim) must equal attention_head_dim"
        assert sum(conditioning_tensor_rope_axes_dim) == attention_head_dim, \
            "sum(conditioning_tensor_rope_axes_dim) must equal attention_head_dim"
        if conditioning_tensor_intermediate_proj_dim not in (None, num_attention_heads * attention_head_dim):
            raise NotImplementedError("conditioning_tensor_intermediate_proj_dim must equal the hidden size")
        self._hidden = num_attention_heads * attention_head_dim
        self._device = torch.device("cpu")
        self._precision = os.environ.get("LADCAST_B200_PRECISION", "bf16")
        self._handle = None
        self._geometry = None
        self._cached = None  # (known data_ptr, ...) while inside cached_conditioning()
        self._sd: Dict[str, torch.Tensor] = {}  # filled by load_state_dict, or lazily with a random init
        self._loaded = False  # a checkpoint was loaded: never paper over missing tensors with a random init

    def _materialize(self):
        """Random initialisation of a FRESH model only (from_config without a checkpoint).  After load_state_dict the
        missing keys stay missing, so that lc_denoiser_finalize fails loudly ("missing checkpoint tensor")."""
        if self._loaded:
            return
        for k, shp in self.param_shapes().items():
            if k not in self._sd:
                self._sd[k] = self._init_param(k, shp)

    # ------------------------------------------------------------------ parameters / checkpoint surface
    @staticmethod
    def _init_param(key, shape):
        if key.endswith(".weight") and len(shape) == 1:
            return torch.ones(shape)  # norm scales
        if key.endswith(".weight"):
            return _linear_init(shape)
        if key.rsplit(".", 2)[-2].startswith("norm") and len(key.split(".")) > 2 and "linear" not in key:
            return torch.zeros(shape)  # LayerNorm biases
        return (torch.rand(shape) * 2 - 1) * 0.02

    def param_shapes(self) -> Dict[str, Tuple[int, ...]]:
        c = self.config
        d, hd, mlp = self._hidden, c.attention_head_dim, int(self._hidden * c.mlp_ratio)
        s: Dict[str, Tuple[int, ...]] = {}

        def lin(n, o, i):
            s[n + ".weight"], s[n + ".bias"] = (o, i), (o,)

        s["x_embedder.proj.weight"], s["x_embedder.proj.bias"] = (d, c.in_channels, 1, 1, 1), (d,)
        s["context_embedder.proj.weight"] = (d, c.conditioning_tensor_in_channels, 1, 1, 1)
        s["context_embedder.proj.bias"] = (d,)
        for pre in ("context_refiner.time_text_embed", "time_text_embed"):
            lin(pre + ".timestep_embedder.linear_1", d, 256)
            lin(pre + ".timestep_embedder.linear_2", d, d)
            lin(pre + ".text_embedder.linear_1", d, d)
            lin(pre + ".text_embedder.linear_2", d, d)
        lin("context_refiner.proj_in", d, d)
        for i in range(c.num_refiner_layers):
            p = f"context_refiner.token_refiner.refiner_blocks.{i}"
            for n in ("norm1", "norm2"):
                s[f"{p}.{n}.weight"], s[f"{p}.{n}.bias"] = (d,), (d,)
            for n in ("to_q", "to_k", "to_v"):
                lin(f"{p}.attn.{n}", d, d)
            s[f"{p}.attn.norm_q.weight"], s[f"{p}.attn.norm_k.weight"] = (hd,), (hd,)
            lin(f"{p}.ff.net.0.proj", mlp, d)
            lin(f"{p}.ff.net.2", d, mlp)
            lin(f"{p}.norm_out.linear", 2 * d, d)
        if c.incl_time_elapsed:
            lin("time_elapsed_embed.linear_1", 2 * d, 256)
            lin("time_elapsed_embed.linear_2", 2 * d, 2 * d)
        for i in range(c.num_layers):
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
        for i in range(c.num_single_layers):
            p = f"single_transformer_blocks.{i}"
            for n in ("to_q", "to_k", "to_v"):
                lin(f"{p}.attn.{n}", d, d)
            s[f"{p}.attn.norm_q.weight"], s[f"{p}.attn.norm_k.weight"] = (hd,), (hd,)
            lin(f"{p}.norm.linear", 3 * d, d)
            lin(f"{p}.proj_mlp", mlp, d)
            lin(f"{p}.proj_out", d, d + mlp)
        lin("norm_out.linear", 2 * d, d)
        lin("proj_out", c.out_channels, d)
        return s

    def state_dict(self):
        self._materialize()
        return dict(self._sd)

    def load_state_dict(self, state_dict, strict: bool = True):
        shapes = self.param_shapes()
        missing = [k for k in shapes if k not in state_dict]
        unexpected = [k for k in state_dict if k not in shapes]
        if strict and (missing or unexpected):
            raise RuntimeError(f"Error(s) in loading state_dict: missing {missing[:5]}, unexpected {unexpected[:5]}")
        for k, shp in shapes.items():
            if k in state_dict:
                t = state_dict[k]
                if tuple(t.shape) != tuple(shp):
                    raise RuntimeError(f"size mismatch for {k}: {tuple(t.shape)} vs {tuple(shp)}")
                self._sd[k] = t.detach().to("cpu", torch.float32).contiguous()
        if not self._loaded:
            # tensors a previous random init put there must not survive next to a (partial) checkpoint
            for k in [k for k in self._sd if k not in state_dict]:
                del self._sd[k]
        self._loaded = True
        if missing:
            import warnings

            warnings.warn(f"LaDCastTransformer3DModel.load_state_dict(strict=False): {len(missing)} tensors are missing "
                          f"(e.g. {missing[:3]}); the model will refuse to run until they are loaded")
        self._release()
        return missing, unexpected

    def parameters(self):
        return iter(self.state_dict().values())

    def num_parameters(self):
        n = 0
        for shp in self.param_shapes().values():
            k = 1
            for v in shp:
                k *= v
            n += k
        return n

    @property
    def dtype(self):
        return torch.float32  # I/O dtype; checkpoint dtype.  Internal compute precision: see set_precision().

    @property
    def device(self):
        return self._device

    def eval(self):
        return self

    def requires_grad_(self, flag=False):
        return self

    def to(self, *args, **kwargs):
        for a in list(args) + list(kwargs.values()):
            if isinstance(a, (str, torch.device)):
                dev = torch.device(a)
                if dev.type == "cuda" and dev.index is None:
                    dev = torch.device("cuda", torch.cuda.current_device())
                if dev != self._device:
                    self._release()
                    self._device = dev
        return self

    def cuda(self, device=None):
        return self.to(torch.device("cuda", device if device is not None else torch.cuda.current_device()))

    def set_precision(self, precision: str):
        """'bf16' (tcgen05 tensor-core path, default) or 'fp32' (SIMT validation path)."""
        assert precision in ("bf16", "fp32")
        if precision != self._precision:
            self._release()
            self._precision = precision
        return self

    # ------------------------------------------------------------------ native handle
    def _release(self):
        if self._handle is not None:
            _lib.load().lc_denoiser_destroy(self._handle)
        self._handle = None
        self._geometry = None
        self._cached = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _ensure_handle(self):
        if self._handle is not None:
            return
        if self._device.type != "cuda":
            raise _lib.LadcastB200Error("LaDCastTransformer3DModel runs on CUDA only (sm_100a); call .to('cuda') first")
        lib = _lib.load()
        c = self.config
        cfg = _lib.DenoiserCfg(c.in_channels, c.out_channels, c.conditioning_tensor_in_channels, c.num_attention_heads,
                               c.attention_head_dim, c.num_layers, c.num_single_layers, c.num_refiner_layers,
                               int(self._hidden * c.mlp_ratio), int(bool(c.incl_time_elapsed)),
                               _lib.PRECISION_F32 if self._precision == "fp32" else _lib.PRECISION_BF16)
        h = ctypes.c_void_p()
        self._materialize()
        _lib.check(lib.lc_denoiser_create(ctypes.byref(cfg), ctypes.byref(h)), "lc_denoiser_create")
        with torch.cuda.device(self._device):
            st = _lib.stream()
            for k, v in self._sd.items():
                dv = v.to(self._device, torch.float32).contiguous()
                shp = (ctypes.c_int64 * dv.dim())(*dv.shape)
                _lib.check(lib.lc_denoiser_load(h, k.encode(), _lib.ptr(dv), shp, dv.dim(), st), f"lc_denoiser_load({k})")
                del dv
            _lib.check(lib.lc_denoiser_finalize(h, st), "lc_denoiser_finalize")
        self._handle = h

    def _ensure_geometry(self, batch, t_in, t_out, height, width):
        geo = (t_in, t_out, height, width)
        if self._geometry is not None and self._geometry[0] == geo and self._geometry[1] >= batch:
            return
        lib = _lib.load()
        (cp, sp), (cc, sc) = rope_tables(self.config, t_in, t_out, height, width)
        dev = [t.to(self._device) for t in (cp, sp, cc, sc)]
        max_b = max(batch, self._geometry[1] if self._geometry and self._geometry[0] == geo else 0)
        _lib.check(lib.lc_denoiser_set_geometry(self._handle, max_b, t_in, t_out, height, width, *[_lib.ptr(t) for t in dev],
                                                _lib.stream()), "lc_denoiser_set_geometry")
        torch.cuda.current_stream().synchronize()
        self._geometry = (geo, max_b)

    def prepare(self, conditioning_tensors: torch.Tensor, time_elapsed=None, t_out: int = 1):
        """Step-invariant part of an AR step (context embedding, date MLP).  `forward` calls it itself unless
        invoked inside `cached_conditioning`."""
        self._ensure_handle()
        lib = _lib.load()
        known = conditioning_tensors.to(self._device, torch.float32).contiguous()
        B, _, t_in, H, W = known.shape
        with torch.cuda.device(self._device):
            self._ensure_geometry(B, t_in, t_out, H, W)
            year, n_ts = None, 0
            if time_elapsed is not None and self.config.incl_time_elapsed:
                stamps = time_elapsed.reshape(-1).tolist() if isinstance(time_elapsed, torch.Tensor) else list(time_elapsed)
                year = year_sincos_embedding(stamps).to(self._device)
                n_ts = len(stamps)
            _lib.check(lib.lc_denoiser_prepare(self._handle, _lib.ptr(known), B, _lib.ptr(year), n_ts, _lib.stream()),
                       "lc_denoiser_prepare")
        return known

    @contextmanager
    def cached_conditioning(self, conditioning_tensors, time_elapsed=None, t_out: int = 1):
        """Hoists `prepare` out of a denoising loop: every `forward` inside the block reuses the cached context."""
        self.prepare(conditioning_tensors, time_elapsed, t_out)
        self._cached = self._cache_key(conditioning_tensors, time_elapsed, t_out)
        try:
            yield self
        finally:
            self._cached = None

    @staticmethod
    def _cache_key(conditioning_tensors, time_elapsed, t_out):
        """Identity of a prepared context: the conditioning tensor (storage, version, shape), the dates and T_out.  A
        forward with anything else re-runs `prepare` instead of silently reusing a stale context."""
        stamps = None
        if time_elapsed is not None:
            stamps = tuple(int(v) for v in (time_elapsed.reshape(-1).tolist() if isinstance(time_elapsed, torch.Tensor)
                                            else time_elapsed))
        return (conditioning_tensors.data_ptr(), conditioning_tensors._version, tuple(conditioning_tensors.shape),
                str(conditioning_tensors.device), stamps, int(t_out))

    # ------------------------------------------------------------------ forward
    def forward(self, hidden_states: torch.Tensor, timestep: torch.Tensor, conditioning_tensors: torch.Tensor,
                time_elapsed: Optional[torch.LongTensor] = None, attention_kwargs: Optional[Dict[str, Any]] = None,
                return_dict: bool = True, coords: Optional[torch.Tensor] = None):
        B, c_in, t_out, H, W = hidden_states.shape
        if conditioning_tensors.shape[0] != B:
            raise ValueError("conditioning_tensors and hidden_states must have the same batch size")
        if c_in != self.config.in_channels:
            raise ValueError(f"hidden_states has {c_in} channels, the model expects in_channels={self.config.in_channels}")
        if tuple(conditioning_tensors.shape[-2:]) != (H, W):
            raise ValueError("conditioning_tensors and hidden_states must share the spatial size")
        if self._cached is None or self._cached != self._cache_key(conditioning_tensors, time_elapsed, t_out):
            self.prepare(conditioning_tensors, time_elapsed, t_out)
        if self._geometry is None or self._geometry[0] != (conditioning_tensors.shape[2], t_out, H, W):
            raise _lib.LadcastB200Error("hidden_states geometry differs from the prepared (T_in, T_out, H, W)")
        lib = _lib.load()
        with torch.cuda.device(self._device):
            x = hidden_states.to(self._device, torch.float32).contiguous()
            t = timestep.to(self._device, torch.float32).reshape(-1).contiguous()
            if t.numel() not in (1, B):
                raise ValueError("timestep must have 1 or batch_size entries")
            out = torch.empty((B, self.config.out_channels, t_out, H, W), device=self._device, dtype=torch.float32)
            _lib.check(lib.lc_denoiser_forward(self._handle, _lib.ptr(x), _lib.ptr(t), t.numel(), _lib.ptr(out), _lib.stream()),
                       "lc_denoiser_forward")
        if not return_dict:
            return (out,)
        return Transformer2DModelOutput(sample=out)

    __call__ = forward

    def debug_read(self, name: str, numel: int) -> torch.Tensor:
        out = torch.empty(numel, device=self._device, dtype=torch.float32)
        _lib.check(_lib.load().lc_denoiser_debug_read(self._handle, name.encode(), _lib.ptr(out), numel, _lib.stream()),
                   "lc_denoiser_debug_read")
        return out
