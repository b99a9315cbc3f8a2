"""Shared plumbing of the drop-in model classes: config object, diffusers-format checkpoint I/O
(`<dir>/config.json` + `<dir>/diffusion_pytorch_model.safetensors`, fp32, reference key names)."""
import inspect
import json
import os

import torch

WEIGHTS_NAME = "diffusion_pytorch_model.safetensors"
CONFIG_NAME = "config.json"


class Config(dict):
    """Key- and attribute-accessible, like diffusers' FrozenDict."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


def capture_config(cls, kwargs):
    sig = inspect.signature(cls.__init__).parameters
    cfg = {k: p.default for k, p in sig.items() if k != "self"}
    cfg.update(kwargs)
    return Config(cfg)


class CheckpointMixin:
    """from_config / from_pretrained / save_pretrained / state_dict surface used by the reference's callers
    (evaluate/pred_rollout.py:298-331)."""

    _class_name = ""

    @classmethod
    def from_config(cls, config, **kwargs):
        sig = inspect.signature(cls.__init__).parameters
        init = {k: v for k, v in dict(config).items() if k in sig and not k.startswith("_")}
        init.update({k: v for k, v in kwargs.items() if k in sig})
        return cls(**init)

    @classmethod
    def from_pretrained(cls, path, subfolder=None, **kwargs):
        from safetensors.torch import load_file

        if subfolder:
            path = os.path.join(path, subfolder)
        with open(os.path.join(path, CONFIG_NAME)) as f:
            model = cls.from_config(json.load(f))
        model.load_state_dict(load_file(os.path.join(path, WEIGHTS_NAME)), strict=True)
        return model.eval()

    def save_pretrained(self, save_directory, **kwargs):
        from safetensors.torch import save_file

        os.makedirs(save_directory, exist_ok=True)
        cfg = dict(self.config)
        cfg["_class_name"] = self._class_name or type(self).__name__
        cfg["_diffusers_version"] = "0.32.1"
        with open(os.path.join(save_directory, CONFIG_NAME), "w") as f:
            json.dump(cfg, f, indent=2, sort_keys=True, default=list)
        sd = {k: v.detach().to("cpu", torch.float32).contiguous() for k, v in self.state_dict().items()}
        save_file(sd, os.path.join(save_directory, WEIGHTS_NAME), metadata={"format": "pt"})
