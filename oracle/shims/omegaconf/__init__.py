"""ORACLE SUPPORT (test infrastructure): the four OmegaConf calls /root/reference/setup.py:69-89 makes, for running the
UNMODIFIED reference train.py in this image, where omegaconf is absent.  Backed by the product's own loader
(motion324_b200/utils/config.py), which restates the same semantics."""
import yaml

from motion324_b200.utils import config as _cfg


class OmegaConf:
    @staticmethod
    def load(path):
        with open(path) as f:
            return yaml.safe_load(f) or {}

    @staticmethod
    def from_cli(args):
        return _cfg.apply_overrides({}, list(args))

    @staticmethod
    def merge(a, b):
        def rec(x, y):
            out = dict(x)
            for k, v in y.items():
                out[k] = rec(out[k], v) if isinstance(v, dict) and isinstance(out.get(k), dict) else v
            return out
        return rec(a, b)

    @staticmethod
    def to_container(cfg, resolve=True):
        return _cfg._resolve(cfg) if resolve else cfg
