"""ModelMixin subset: dtype/device, from_config (via ConfigMixin), save/from_pretrained in the
diffusers on-disk layout (config.json + diffusion_pytorch_model.safetensors)."""
import os

import torch
import torch.nn as nn

WEIGHTS_NAME = "diffusion_pytorch_model.safetensors"


class ModelMixin(nn.Module):
    config_name = "config.json"
    _supports_gradient_checkpointing = False

    def __init__(self):
        super().__init__()

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    def enable_gradient_checkpointing(self):
        self.gradient_checkpointing = True

    def save_pretrained(self, save_directory, **kwargs):
        from safetensors.torch import save_file

        os.makedirs(save_directory, exist_ok=True)
        self.save_config(save_directory)
        sd = {k: v.detach().contiguous().cpu() for k, v in self.state_dict().items()}
        save_file(sd, os.path.join(save_directory, WEIGHTS_NAME), metadata={"format": "pt"})

    @classmethod
    def from_pretrained(cls, path, subfolder=None, torch_dtype=None, **kwargs):
        from safetensors.torch import load_file

        if subfolder:
            path = os.path.join(path, subfolder)
        cfg = cls.load_config(path)
        model = cls.from_config(cfg)
        model.load_state_dict(load_file(os.path.join(path, WEIGHTS_NAME)), strict=True)
        if torch_dtype is not None:
            model = model.to(torch_dtype)
        return model.eval()
