import logging as _pylogging
from collections import OrderedDict
from dataclasses import fields, is_dataclass

import torch
from packaging import version as _v

USE_PEFT_BACKEND = False


class _Logging:
    @staticmethod
    def get_logger(name):
        return _pylogging.getLogger(name)


logging = _Logging()


def is_torch_version(op, ver):
    import operator

    ops = {">": operator.gt, ">=": operator.ge, "==": operator.eq, "<": operator.lt, "<=": operator.le}
    return ops[op](_v.parse(torch.__version__.split("+")[0]), _v.parse(ver))


def scale_lora_layers(model, weight):
    return None


def unscale_lora_layers(model, weight=None):
    return None


class BaseOutput(OrderedDict):
    """dataclass + ordered dict; supports .field, ["field"], [int] (diffusers utils/outputs.py)."""

    def __post_init__(self):
        assert is_dataclass(self)
        for f in fields(self):
            v = getattr(self, f.name)
            if v is not None:
                OrderedDict.__setitem__(self, f.name, v)

    def __getitem__(self, k):
        if isinstance(k, str):
            return dict(self.items())[k]
        return self.to_tuple()[k]

    def __setattr__(self, name, value):
        if name in self.keys() and value is not None:
            OrderedDict.__setitem__(self, name, value)
        super().__setattr__(name, value)

    def __setitem__(self, key, value):
        OrderedDict.__setitem__(self, key, value)
        super().__setattr__(key, value)

    def to_tuple(self):
        return tuple(self[k] for k in self.keys())
