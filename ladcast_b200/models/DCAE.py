"""Drop-in for `ladcast.models.DCAE.AutoencoderDC` (reference models/DCAE.py:735-1087): same constructor keywords,
`.config`, checkpoint layout / key names, `encode(...)` / `decode(...)` signatures.  Both run the sm_100a kernels
behind `lc_dcae_*` (decoder: every AR step; encoder: once per forecast init time)."""
import ctypes
import os
from dataclasses import dataclass
from typing import Dict, Optional, Tuple, Union

import torch

from .. import _lib
from .modeling import CheckpointMixin, capture_config


@dataclass
class DecoderOutput:
    sample: torch.Tensor
    commit_loss: Optional[torch.Tensor] = None

    def __getitem__(self, i):
        return (self.sample,)[i]


@dataclass
class EncoderOutput:
    latent: torch.Tensor


class AutoencoderDC(CheckpointMixin):
    _class_name = "AutoencoderDC"
    MAX_FRAMES_PER_CALL = 80

    def __init__(self, in_channels: int = 3, out_channels: Optional[int] = None, temb_channels: Optional[int] = None,
                 latent_channels: int = 32, attention_head_dim: int = 32,
                 encoder_block_types: Union[str, Tuple[str]] = "ResBlock",
                 decoder_block_types: Union[str, Tuple[str]] = "ResBlock",
                 encoder_block_out_channels: Tuple[int, ...] = (128, 256, 512, 512, 1024, 1024),
                 decoder_block_out_channels: Tuple[int, ...] = (128, 256, 512, 512, 1024, 1024),
                 encoder_layers_per_block: Tuple[int] = (2, 2, 2, 3, 3, 3),
                 decoder_layers_per_block: Tuple[int] = (3, 3, 3, 3, 3, 3),
                 encoder_qkv_multiscales=((), (), (), (5,), (5,), (5,)),
                 decoder_qkv_multiscales=((), (), (), (5,), (5,), (5,)), upsample_block_type: str = "pixel_shuffle",
                 downsample_block_type: str = "pixel_unshuffle", decoder_norm_types: Union[str, Tuple[str]] = "rms_norm",
                 decoder_act_fns: Union[str, Tuple[str]] = "silu", scaling_factor: float = 1.0,
                 static_channels: int = 0) -> None:
        kw = dict(locals())
        kw.pop("self")
        self.config = capture_config(type(self), kw)
        if temb_channels is not None:
            raise NotImplementedError("temb-conditioned DC-AE is not part of the V0.1.X checkpoint")
        if upsample_block_type != "pixel_shuffle" or decoder_norm_types != "rms_norm" or decoder_act_fns != "silu":
            raise NotImplementedError("only the V0.1.X decoder variant (pixel_shuffle / rms_norm / silu) is implemented")
        n = len(decoder_block_out_channels)
        etypes = (encoder_block_types,) * n if isinstance(encoder_block_types, str) else tuple(encoder_block_types)
        # the encoder is optional: configurations the CUDA encoder does not cover only disable encode()
        self._enc_error = None
        if len(encoder_block_out_channels) != n:
            self._enc_error = "encoder and decoder must have the same number of stages"
        elif downsample_block_type != "pixel_unshuffle":
            self._enc_error = "only the pixel_unshuffle down-sampling variant is implemented"
        elif encoder_layers_per_block[0] <= 0:
            self._enc_error = "encoder_layers_per_block[0] == 0 (down-sampling conv_in) is not implemented"
        else:
            for i, t in enumerate(etypes):
                if t == "EfficientViTBlock" and tuple(encoder_qkv_multiscales[i]) != (5,) and encoder_layers_per_block[i] > 0:
                    self._enc_error = "EfficientViT stages must use qkv_multiscales == (5,)"
        self._etypes = etypes
        types = (decoder_block_types,) * n if isinstance(decoder_block_types, str) else tuple(decoder_block_types)
        for i, t in enumerate(types):
            if t == "EfficientViTBlock" and tuple(decoder_qkv_multiscales[i]) != (5,) and decoder_layers_per_block[i] > 0:
                raise NotImplementedError("EfficientViT stages must use qkv_multiscales == (5,)")
        self._types = types
        self.static_channels = static_channels
        self.spatial_compression_ratio = 2 ** (n - 1)
        self.temporal_compression_ratio = 1
        self.use_slicing = False
        self.use_tiling = False
        self._device = torch.device("cpu")
        self._precision = os.environ.get("LADCAST_B200_PRECISION", "bf16")
        self._handle = None
        self._reserved = None
        self._sd: Dict[str, torch.Tensor] = {}
        self._loaded_groups = None  # None = fresh model (random init allowed); else subset of {"decoder", "encoder"}

    # ------------------------------------------------------------------ parameters
    def encoder_param_shapes(self) -> Dict[str, Tuple[int, ...]]:
        """encoder.* keys / shapes (models/DCAE.py:539-615)."""
        if self._enc_error is not None:
            return {}
        c = self.config
        ch, layers, hd = list(c.encoder_block_out_channels), list(c.encoder_layers_per_block), c.attention_head_dim
        s: Dict[str, Tuple[int, ...]] = {"encoder.conv_in.weight": (ch[0], c.in_channels, 3, 3), "encoder.conv_in.bias": (ch[0],)}
        j, n = 0, len(ch)
        for i in range(n):
            for _ in range(layers[i]):
                p, C = f"encoder.down_blocks.{j}", ch[i]
                if self._etypes[i] == "ResBlock":
                    s.update({f"{p}.conv1.weight": (C, C, 3, 3), f"{p}.conv1.bias": (C,), f"{p}.conv2.weight": (C, C, 3, 3),
                              f"{p}.norm.weight": (C,), f"{p}.norm.bias": (C,)})
                else:
                    inner = (C // hd) * hd
                    for nm in ("to_q", "to_k", "to_v"):
                        s[f"{p}.attn.{nm}.weight"] = (inner, C)
                    s.update({f"{p}.attn.to_qkv_multiscale.0.proj_in.weight": (3 * inner, 1, 5, 5),
                              f"{p}.attn.to_qkv_multiscale.0.proj_out.weight": (3 * inner, hd, 1, 1),
                              f"{p}.attn.to_out.weight": (C, 2 * inner), f"{p}.attn.norm_out.weight": (C,),
                              f"{p}.attn.norm_out.bias": (C,), f"{p}.conv_out.conv_inverted.weight": (8 * C, C, 1, 1),
                              f"{p}.conv_out.conv_inverted.bias": (8 * C,), f"{p}.conv_out.conv_depth.weight": (8 * C, 1, 3, 3),
                              f"{p}.conv_out.conv_depth.bias": (8 * C,), f"{p}.conv_out.conv_point.weight": (C, 4 * C, 1, 1),
                              f"{p}.conv_out.norm.weight": (C,), f"{p}.conv_out.norm.bias": (C,)})
                j += 1
            if i < n - 1 and layers[i] > 0:
                s[f"encoder.down_blocks.{j}.conv.weight"] = (ch[i + 1] // 4, ch[i], 3, 3)
                s[f"encoder.down_blocks.{j}.conv.bias"] = (ch[i + 1] // 4,)
                j += 1
        s["encoder.conv_out.weight"], s["encoder.conv_out.bias"] = (c.latent_channels, ch[-1], 3, 3), (c.latent_channels,)
        return s

    def param_shapes(self) -> Dict[str, Tuple[int, ...]]:
        s = dict(self.encoder_param_shapes())
        s.update(self.decoder_param_shapes())
        return s

    def decoder_param_shapes(self) -> Dict[str, Tuple[int, ...]]:
        c = self.config
        ch, layers, hd = list(c.decoder_block_out_channels), list(c.decoder_layers_per_block), c.attention_head_dim
        oc = c.out_channels if c.out_channels is not None else c.in_channels
        s: Dict[str, Tuple[int, ...]] = {"decoder.conv_in.weight": (ch[-1], c.latent_channels, 3, 3),
                                         "decoder.conv_in.bias": (ch[-1],)}
        j, n = 0, len(ch)
        for i in reversed(range(n)):
            if i < n - 1 and layers[i] > 0:
                s[f"decoder.up_blocks.{j}.conv.weight"] = (4 * ch[i], ch[i + 1], 3, 3)
                s[f"decoder.up_blocks.{j}.conv.bias"] = (4 * ch[i],)
                j += 1
            for _ in range(layers[i]):
                p, C = f"decoder.up_blocks.{j}", ch[i]
                if self._types[i] == "ResBlock":
                    s.update({f"{p}.conv1.weight": (C, C, 3, 3), f"{p}.conv1.bias": (C,), f"{p}.conv2.weight": (C, C, 3, 3),
                              f"{p}.norm.weight": (C,), f"{p}.norm.bias": (C,)})
                else:
                    inner = (C // hd) * hd
                    for nm in ("to_q", "to_k", "to_v"):
                        s[f"{p}.attn.{nm}.weight"] = (inner, C)
                    s.update({f"{p}.attn.to_qkv_multiscale.0.proj_in.weight": (3 * inner, 1, 5, 5),
                              f"{p}.attn.to_qkv_multiscale.0.proj_out.weight": (3 * inner, hd, 1, 1),
                              f"{p}.attn.to_out.weight": (C, 2 * inner), f"{p}.attn.norm_out.weight": (C,),
                              f"{p}.attn.norm_out.bias": (C,), f"{p}.conv_out.conv_inverted.weight": (8 * C, C, 1, 1),
                              f"{p}.conv_out.conv_inverted.bias": (8 * C,), f"{p}.conv_out.conv_depth.weight": (8 * C, 1, 3, 3),
                              f"{p}.conv_out.conv_depth.bias": (8 * C,), f"{p}.conv_out.conv_point.weight": (C, 4 * C, 1, 1),
                              f"{p}.conv_out.norm.weight": (C,), f"{p}.conv_out.norm.bias": (C,)})
                j += 1
        s["decoder.norm_out.weight"], s["decoder.norm_out.bias"] = (ch[0],), (ch[0],)
        s["decoder.conv_out.weight"], s["decoder.conv_out.bias"] = (oc, ch[0], 3, 3), (oc,)
        return s

    def _active_shapes(self) -> Dict[str, Tuple[int, ...]]:
        """Shapes of the sub-modules this instance holds: everything for a fresh model, otherwise only the groups a
        checkpoint provided (a decoder-only load must NOT grow a random encoder: lc_dcae_encode then fails loudly with
        'no encoder.* weights were loaded into this handle')."""
        if self._loaded_groups is None:
            return self.param_shapes()
        s: Dict[str, Tuple[int, ...]] = {}
        if "encoder" in self._loaded_groups:
            s.update(self.encoder_param_shapes())
        if "decoder" in self._loaded_groups:
            s.update(self.decoder_param_shapes())
        return s

    def _materialize(self):
        if self._loaded_groups is not None:
            return
        for k, shp in self.param_shapes().items():
            if k not in self._sd:
                if k.endswith(".weight") and len(shp) > 1:
                    fan = 1
                    for v in shp[1:]:
                        fan *= v
                    self._sd[k] = (torch.rand(shp) * 2 - 1) / fan**0.5
                elif k.endswith(".weight"):
                    self._sd[k] = torch.ones(shp)
                else:
                    self._sd[k] = torch.zeros(shp)

    def state_dict(self):
        self._materialize()
        return dict(self._sd)

    def load_state_dict(self, state_dict, strict: bool = True):
        """`decoder.*` / `encoder.*` tensors (reference key names) are re-packed by lc_dcae_load at first use."""
        shapes = self.param_shapes()
        missing = [k for k in shapes if k not in state_dict]
        if strict:
            # strict = every sub-module (decoder / encoder) that appears in the state dict must be complete, and at
            # least one must appear; a decoder-only (rollout) or encoder-only (compression) checkpoint is legal
            seen = False
            for grp in (self.decoder_param_shapes(), self.encoder_param_shapes()):
                present = [k for k in grp if k in state_dict]
                seen = seen or bool(present)
                if present and len(present) != len(grp):
                    raise RuntimeError(f"Error(s) in loading state_dict: missing {[k for k in grp if k not in state_dict][:5]}")
            if not seen:
                raise RuntimeError(f"Error(s) in loading state_dict: missing {missing[:5]}")
        if self._loaded_groups is None:
            self._sd = {}  # drop a previous random init: only checkpoint tensors may be uploaded from now on
            self._loaded_groups = set()
        for k, t in state_dict.items():
            if k in shapes and tuple(t.shape) != tuple(shapes[k]):
                raise RuntimeError(f"size mismatch for {k}: {tuple(t.shape)} vs {tuple(shapes[k])}")
            if k in shapes or k.startswith("encoder."):
                self._sd[k] = t.detach().to("cpu", torch.float32).contiguous()
            elif strict:
                raise RuntimeError(f"unexpected key {k}")
        for name, grp in (("decoder", self.decoder_param_shapes()), ("encoder", self.encoder_param_shapes())):
            if grp and all(k in self._sd for k in grp):
                self._loaded_groups.add(name)
        if not self._loaded_groups:
            raise RuntimeError(f"Error(s) in loading state_dict: no complete decoder.* / encoder.* group; missing {missing[:5]}")
        self._release()
        return missing, []

    def parameters(self):
        return iter(self.state_dict().values())

    @property
    def dtype(self):
        return torch.float32

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

    def set_precision(self, precision: str):
        assert precision in ("bf16", "fp32")
        if precision != self._precision:
            self._release()
            self._precision = precision
        return self

    # ------------------------------------------------------------------ native handle
    def _release(self):
        if self._handle is not None:
            _lib.load().lc_dcae_destroy(self._handle)
        self._handle, self._reserved = None, None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _ensure_handle(self):
        if self._handle is not None:
            return
        if self._device.type != "cuda":
            raise _lib.LadcastB200Error("AutoencoderDC runs on CUDA only (sm_100a); call .to('cuda') first")
        lib = _lib.load()
        c = self.config
        ch, layers = list(c.decoder_block_out_channels), list(c.decoder_layers_per_block)
        cfg = _lib.DcaeCfg()
        cfg.latent_channels = c.latent_channels
        cfg.out_channels = c.out_channels if c.out_channels is not None else c.in_channels
        cfg.head_dim = c.attention_head_dim
        cfg.n_stages = len(ch)
        cfg.precision = _lib.PRECISION_F32 if self._precision == "fp32" else _lib.PRECISION_BF16
        for i in range(len(ch)):
            cfg.stage_channels[i], cfg.stage_layers[i] = ch[i], layers[i]
            cfg.stage_is_evit[i] = int(self._types[i] == "EfficientViTBlock")
        if self._enc_error is None:
            cfg.in_channels = c.in_channels
            for i in range(len(ch)):
                cfg.enc_stage_channels[i] = c.encoder_block_out_channels[i]
                cfg.enc_stage_layers[i] = c.encoder_layers_per_block[i]
                cfg.enc_stage_is_evit[i] = int(self._etypes[i] == "EfficientViTBlock")
        h = ctypes.c_void_p()
        self._materialize()
        _lib.check(lib.lc_dcae_create(ctypes.byref(cfg), ctypes.byref(h)), "lc_dcae_create")
        with torch.cuda.device(self._device):
            st = _lib.stream()
            for k in self._active_shapes():
                dv = self._sd[k].to(self._device, torch.float32).contiguous()
                shp = (ctypes.c_int64 * dv.dim())(*dv.shape)
                _lib.check(lib.lc_dcae_load(h, k.encode(), _lib.ptr(dv), shp, dv.dim(), st), f"lc_dcae_load({k})")
                del dv
            _lib.check(lib.lc_dcae_finalize(h, st), "lc_dcae_finalize")
        self._handle = h

    def _decode_native(self, z, keep, mean=None, std=None):
        self._ensure_handle()
        lib = _lib.load()
        z = z.to(self._device, torch.float32).contiguous()
        n, _, h, w = z.shape
        r = self.spatial_compression_ratio
        out = torch.empty(n, keep, h * r, w * r, device=self._device, dtype=torch.float32)
        with torch.cuda.device(self._device):
            chunk = min(n, self.MAX_FRAMES_PER_CALL)
            self._reserve(lib, chunk, h, w)
            mean_d = mean.to(self._device, torch.float32).contiguous() if mean is not None else None
            std_d = std.to(self._device, torch.float32).contiguous() if std is not None else None
            for i in range(0, n, chunk):
                m = min(chunk, n - i)
                _lib.check(lib.lc_dcae_decode(self._handle, _lib.ptr(z[i : i + m]), m, h, w, _lib.ptr(out[i : i + m]), keep,
                                              _lib.ptr(mean_d), _lib.ptr(std_d), _lib.stream()), "lc_dcae_decode")
        return out

    # ------------------------------------------------------------------ reference API
    def decode(self, z: torch.Tensor, return_dict: bool = True, temb: Optional[torch.Tensor] = None, embedded_t: bool = False,
               return_static=False):
        if temb is not None:
            raise NotImplementedError("temb-conditioned decoding is not part of the V0.1.X checkpoint")
        oc = self.config.out_channels if self.config.out_channels is not None else self.config.in_channels
        keep = oc if (return_static or not self.static_channels) else oc - self.static_channels
        decoded = self._decode_native(z, keep)
        if not return_dict:
            return (decoded,)
        return DecoderOutput(sample=decoded)

    def decode_fused(self, z: torch.Tensor, mean: Optional[torch.Tensor], std: Optional[torch.Tensor]) -> torch.Tensor:
        """decode + (x*std + mean) in the last kernel's epilogue — what decode_latent_ens needs."""
        oc = self.config.out_channels if self.config.out_channels is not None else self.config.in_channels
        keep = oc - self.static_channels if self.static_channels else oc
        return self._decode_native(z, keep, mean, std)

    def decode_ens_fused(self, latents: torch.Tensor, mean: Optional[torch.Tensor] = None,
                         std: Optional[torch.Tensor] = None, extract_first: Optional[int] = None,
                         latent_mean: Optional[torch.Tensor] = None, latent_std: Optional[torch.Tensor] = None,
                         target_std: float = 0.5, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """`decode_latent_ens` (reference pipelines/utils.py:52-80) in one native pass: latents (B, C, T, h, w) are read
        in place (no permute / reshape copies), every (b, t < extract_first) frame is decoded and written straight to
        out (B, 84, extract_first, 8h, 8w), de-normalised with mean / std in the last epilogue.  With latent_mean /
        latent_std the latents are first mapped back from normalised units, (z / target_std) * std + mean, inside the
        first kernel (what roll_out_serial does before decoding, pipelines/utils.py:571-577)."""
        self._ensure_handle()
        lib = _lib.load()
        z = latents.to(self._device, torch.float32).contiguous()
        B, C, T, h, w = z.shape
        take = T if extract_first is None else int(extract_first)
        oc = self.config.out_channels if self.config.out_channels is not None else self.config.in_channels
        keep = oc - self.static_channels if self.static_channels else oc
        r = self.spatial_compression_ratio
        if out is None:
            out = torch.empty(B, keep, take, h * r, w * r, device=self._device, dtype=torch.float32)
        elif tuple(out.shape) != (B, keep, take, h * r, w * r) or not out.is_contiguous() or out.dtype != torch.float32:
            raise ValueError("out must be a contiguous float32 tensor of shape (B, keep_channels, extract_first, 8h, 8w)")
        dev = lambda t: t.to(self._device, torch.float32).contiguous() if t is not None else None  # noqa: E731
        mean_d, std_d, lm_d, ls_d = dev(mean), dev(std), dev(latent_mean), dev(latent_std)
        n = B * take
        with torch.cuda.device(self._device):
            chunk = min(n, self.MAX_FRAMES_PER_CALL)
            self._reserve(lib, chunk, h, w)
            for f0 in range(0, n, chunk):
                m = min(chunk, n - f0)
                _lib.check(lib.lc_dcae_decode_ens(self._handle, _lib.ptr(z), B, T, take, f0, m, h, w, _lib.ptr(out), keep,
                                                  _lib.ptr(mean_d), _lib.ptr(std_d), _lib.ptr(lm_d), _lib.ptr(ls_d),
                                                  float(target_std), _lib.stream()), "lc_dcae_decode_ens")
        return out

    def _reserve(self, lib, chunk, h, w):
        if self._reserved is None or self._reserved[0] < chunk or self._reserved[1:] != (h, w):
            _lib.check(lib.lc_dcae_reserve(self._handle, chunk, h, w, _lib.stream()), "lc_dcae_reserve")
            self._reserved = (chunk, h, w)

    def _encode_native(self, x, mean=None, std=None, target_std=0.5):
        if self._enc_error is not None:
            raise NotImplementedError(f"AutoencoderDC.encode: {self._enc_error}")
        self._ensure_handle()
        lib = _lib.load()
        x = x.to(self._device, torch.float32).contiguous()
        n, cin, H, W = x.shape
        r = self.spatial_compression_ratio
        if cin != self.config.in_channels:
            raise ValueError(f"encode expects {self.config.in_channels} channels (fields + static), got {cin}")
        if H % r or W % r:
            raise ValueError(f"field size {H}x{W} is not a multiple of the compression ratio {r}")
        h, w = H // r, W // r
        out = torch.empty(n, self.config.latent_channels, h, w, device=self._device, dtype=torch.float32)
        with torch.cuda.device(self._device):
            chunk = min(n, self.MAX_FRAMES_PER_CALL)
            self._reserve(lib, chunk, h, w)
            mean_d = mean.to(self._device, torch.float32).contiguous() if mean is not None else None
            std_d = std.to(self._device, torch.float32).contiguous() if std is not None else None
            for i in range(0, n, chunk):
                m = min(chunk, n - i)
                _lib.check(lib.lc_dcae_encode(self._handle, _lib.ptr(x[i : i + m]), m, H, W, _lib.ptr(out[i : i + m]),
                                              _lib.ptr(mean_d), _lib.ptr(std_d), float(target_std), _lib.stream()),
                           "lc_dcae_encode")
        return out

    def encode(self, x, return_dict: bool = True, temb=None, embedded_t: bool = False, static_conditioning_tensor=None):
        """AutoencoderDC.encode (DCAE.py:964-1000): x [n, C, H, W] (+ static_conditioning_tensor [n, C_s, H, W])."""
        if temb is not None:
            raise NotImplementedError("temb-conditioned encoding is not part of the V0.1.X checkpoint")
        if static_conditioning_tensor is not None:
            x = torch.cat((x.to(self._device), static_conditioning_tensor.to(self._device)), dim=1)
        encoded = self._encode_native(x)
        if not return_dict:
            return (encoded,)
        return EncoderOutput(latent=encoded)

    def encode_fused(self, x, mean, std, target_std: float = 0.5, static_conditioning_tensor=None):
        """encode + normalize_transform_3D ((z - mean) / std * target_std) in the last kernel."""
        if static_conditioning_tensor is not None:
            x = torch.cat((x.to(self._device), static_conditioning_tensor.to(self._device)), dim=1)
        return self._encode_native(x, mean, std, target_std)

    def forward(self, sample: torch.Tensor, return_dict: bool = True, static_conditioning_tensor=None):
        """encode -> decode (DCAE.py:1058-1087)."""
        z = self.encode(sample, return_dict=False, static_conditioning_tensor=static_conditioning_tensor)[0]
        return self.decode(z, return_dict=return_dict)
