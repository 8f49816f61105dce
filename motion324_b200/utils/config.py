"""configs/dyscene.yaml of the reference as an EasyDict (the keys Motion_Latent_Model reads: Pcd_motion.py:272-287,
309, 351-354, 369, 457, 463, 512-513; model/loss.py:18-22)."""
from .easydict import EasyDict


def make_config(frames=12, drop_rate=0.1, coord_mse_loss_weight=1.0, **training_overrides):
    cfg = EasyDict({
        "model": {
            "class_name": "motion324_b200.model.Pcd_motion.Motion_Latent_Model",
            "feat_dim": 768, "tokens": 64, "pcd_layers": 4,
            "video_encoder": {
                "image_tokenizer": {"image_size": 224, "patch_size": 14, "patch_length": 1, "in_channels": 3},
                "transformer": {"d": 768, "d_head": 64, "n_layer": 16, "special_init": True, "depth_init": True,
                                "use_qk_norm": True, "drop_rate": drop_rate},
            },
        },
        "training": {"frames": frames, "use_checkpoint": True, "grad_checkpoint_every": 1,
                     "coord_mse_loss_weight": coord_mse_loss_weight, "amp_dtype": "bf16", "use_amp": True,
                     "use_tf32": True, "num_shape_samples": 4096, "num_pcd_samples": 4096, "batch_size_per_gpu": 16},
    })
    for k, v in training_overrides.items():
        cfg.training[k] = v
    return cfg


# --------------------------------------------------------------------------------------------- CLI config loader
# Same behaviour as /root/reference/setup.py:52-89 (YAML file + "key=value" overrides with dotted keys, whitespace around '='
# repaired, new top-level keys accepted, values typed like YAML scalars) without omegaconf, which is absent in this image.

def process_overrides(overrides):
    """setup.py:52-67: 'a = 1' / 'a= 1' -> 'a=1'."""
    import re
    fixed = re.sub(r"(\S+)\s*=\s*(\S+)", r"\1=\2", " ".join(overrides))
    return re.findall(r"[^\s=]+=\S+|\S+", fixed)


def apply_overrides(cfg, overrides):
    import yaml
    for item in process_overrides(overrides):
        if "=" not in item:
            raise ValueError(f"override {item!r} is not of the form key=value")
        key, raw = item.split("=", 1)
        value = _scalar(yaml.safe_load(raw)) if raw != "" else None
        node = cfg
        parts = key.split(".")
        for p in parts[:-1]:
            if p not in node or not isinstance(node[p], dict):
                node[p] = {}
            node = node[p]
        node[parts[-1]] = value
    return cfg


def _scalar(v):
    """YAML 1.1 reads '1e-4' as a string; OmegaConf reads a float."""
    if isinstance(v, str):
        try:
            return float(v) if any(c in v for c in ".eE") and not v.strip().isalpha() else v
        except ValueError:
            return v
    return v


def _resolve(cfg):
    """``${a.b.c}`` interpolation (OmegaConf.to_container(resolve=True)): configs/dyscene.yaml uses it for checkpoint_dir."""
    import re

    def lookup(path):
        node = cfg
        for p in path.split("."):
            node = node[p]
        return node

    def walk(node):
        for k, v in (node.items() if isinstance(node, dict) else enumerate(node)):
            if isinstance(v, (dict, list)):
                walk(v)
            elif isinstance(v, str) and "${" in v:
                whole = re.fullmatch(r"\$\{([^}]+)\}", v)
                node[k] = lookup(whole.group(1)) if whole else re.sub(r"\$\{([^}]+)\}", lambda m: str(lookup(m.group(1))), v)
    walk(cfg)
    return cfg


def load_config(path, overrides=()):
    import yaml
    with open(path) as f:
        cfg = yaml.safe_load(f) or {}
    return EasyDict(_resolve(apply_overrides(cfg, list(overrides))))


def init_config(argv=None):
    """setup.py:69-89: ``--config X.yaml key=value ...`` -> EasyDict."""
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", "-c", required=True)
    ap.add_argument("overrides", nargs="*")
    args = ap.parse_args(argv)
    return load_config(args.config, args.overrides)
