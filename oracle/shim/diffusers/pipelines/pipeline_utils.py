import torch


class DiffusionPipeline:
    """register_modules / _execution_device subset (diffusers 0.32.1 pipelines/pipeline_utils.py)."""

    def __init__(self):
        self._component_names = []

    def register_modules(self, **kwargs):
        for name, module in kwargs.items():
            if name not in self._component_names:
                self._component_names.append(name)
            setattr(self, name, module)

    @property
    def components(self):
        return {k: getattr(self, k) for k in self._component_names}

    @property
    def device(self):
        for m in self.components.values():
            if isinstance(m, torch.nn.Module):
                return m.device
        return torch.device("cpu")

    @property
    def _execution_device(self):
        return self.device

    def to(self, *args, **kwargs):
        for m in self.components.values():
            if isinstance(m, torch.nn.Module):
                m.to(*args, **kwargs)
        return self
