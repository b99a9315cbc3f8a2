"""ConfigMixin / register_to_config restated (diffusers 0.32.1 configuration_utils.py)."""
import functools
import inspect
import json
import os
from collections import OrderedDict


class FrozenDict(OrderedDict):
    """Attribute- and key-accessible config dict.  (diffusers' own class is only nominally frozen: its
    `__setitem__` guard never fires because of name mangling, which is what lets `copy.deepcopy(scheduler)`
    at ladcast/pipelines/utils.py:700 work — so item assignment must stay legal here too.)"""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        for k, v in self.items():
            object.__setattr__(self, k, v)

    def __setitem__(self, k, v):
        super().__setitem__(k, v)
        object.__setattr__(self, k, v)


class ConfigMixin:
    config_name = "config.json"
    ignore_for_config = []

    def register_to_config(self, **kwargs):
        kwargs.pop("kwargs", None)
        if not hasattr(self, "_internal_dict"):
            internal = kwargs
        else:
            internal = {**self._internal_dict, **kwargs}
        self._internal_dict = FrozenDict(internal)

    @property
    def config(self):
        return self._internal_dict

    @classmethod
    def from_config(cls, config, **kwargs):
        config = dict(config)
        sig = inspect.signature(cls.__init__).parameters
        init = {k: v for k, v in config.items() if k in sig and not k.startswith("_")}
        init.update({k: v for k, v in kwargs.items() if k in sig})
        return cls(**init)

    def save_config(self, save_directory):
        os.makedirs(save_directory, exist_ok=True)
        d = dict(self.config)
        d["_class_name"] = self.__class__.__name__
        d["_diffusers_version"] = "0.32.1"
        with open(os.path.join(save_directory, self.config_name), "w") as f:
            json.dump(d, f, indent=2, sort_keys=True, default=lambda o: list(o))

    @classmethod
    def load_config(cls, path, subfolder=None, **kwargs):
        if subfolder:
            path = os.path.join(path, subfolder)
        with open(os.path.join(path, cls.config_name)) as f:
            return json.load(f)


def register_to_config(init):
    @functools.wraps(init)
    def inner_init(self, *args, **kwargs):
        init_kwargs = {k: v for k, v in kwargs.items() if not k.startswith("_")}
        config_init_kwargs = {k: v for k, v in kwargs.items() if k.startswith("_")}
        ignore = getattr(self, "ignore_for_config", [])
        new_kwargs = {}
        signature = inspect.signature(init)
        parameters = {
            name: p.default for i, (name, p) in enumerate(signature.parameters.items()) if i > 0 and name not in ignore
        }
        for arg, name in zip(args, parameters.keys()):
            new_kwargs[name] = arg
        new_kwargs.update(
            {k: init_kwargs.get(k, default) for k, default in parameters.items() if k not in ignore and k not in new_kwargs}
        )
        new_kwargs = {**config_init_kwargs, **new_kwargs}
        getattr(self, "register_to_config")(**new_kwargs)
        init(self, *args, **init_kwargs)

    return inner_init
