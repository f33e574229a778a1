"""YAML config access with attribute + mapping semantics (omegaconf is not installed offline; if it is, its
DictConfig objects work unchanged everywhere a Config is accepted)."""
import copy as _copy
import os

import yaml


class Config(dict):
    """dict with attribute access, `.copy()` (deep), `delattr`, `.get` — the subset of OmegaConf that
    processors/model_manager.py:62-121 and processors/vcfprocessor.py:24-61 use."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        for key, v in list(self.items()):
            if isinstance(v, dict) and not isinstance(v, Config):
                self[key] = Config(v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = Config(v) if isinstance(v, dict) and not isinstance(v, Config) else v

    def __delattr__(self, k):
        try:
            del self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def copy(self):
        return Config(_copy.deepcopy(dict(self)))


def load_yaml(path) -> Config:
    with open(path) as f:
        return Config(yaml.safe_load(f))


PACKAGE_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONFIG_DIR = os.path.join(PACKAGE_ROOT, "configs")
VOCAB_DIR = os.path.join(PACKAGE_ROOT, "vocabs")
