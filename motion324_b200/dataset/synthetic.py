"""Synthetic stand-in for the reference's unreleased training set (README.md:97), pluggable through the reference's own seam
``training.dataset_name=motion324_b200.dataset.synthetic.SyntheticDyscene`` (train.py:50-53): same constructor
``Dataset(config.training)``, same per-item schema as dataset/dyscene.py:315-327 (so train.py's ``collate_fn_with_topology``,
dataset/dyscene.py:331-383, batches it unchanged), random content of the right statistics (SURVEY.md 8(d) "Synthetic inputs").

Items are cheap: one pool of random frames / points per worker process, re-indexed per item, so that 8 ranks x 8 workers do not
saturate the host while the GPUs train."""
import torch
from torch.utils.data import Dataset


def _get(cfg, key, default):
    return cfg.get(key, default) if hasattr(cfg, "get") else getattr(cfg, key, default)


class SyntheticDyscene(Dataset):
    def __init__(self, config, pcd_subdir="pcds", transform=None):
        self.frames = int(_get(config, "frames", 12))
        self.n_points = int(_get(config, "num_pcd_samples", 4096))
        self.n_shape = int(_get(config, "num_shape_samples", 4096))
        self.length = int(_get(config, "synthetic_len", 4096))
        self.image = int(_get(config, "synthetic_image_size", 224))
        self.pool = int(_get(config, "synthetic_pool", 4))
        self._cache = None

    def __len__(self):
        return self.length

    def _pool(self):
        if self._cache is None:
            g = torch.Generator().manual_seed(1234)
            T, N, S, H = self.frames, self.n_points, self.n_shape, self.image
            items = []
            for _ in range(self.pool):
                unit = lambda n: torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
                ref = torch.rand(N, 3, generator=g) - 0.5
                items.append(dict(
                    rgb_video=torch.rand(T, H, H, 3, generator=g),
                    ref_shape_pcd=torch.rand(S, 3, generator=g) - 0.5, ref_shape_normals=unit(S), ref_shape_rgbs=torch.rand(S, 3, generator=g),
                    ref_pcd=ref, ref_normal=unit(N), ref_rgb=torch.rand(N, 3, generator=g),
                    point_clouds=ref[None] + 0.05 * torch.randn(T, N, 3, generator=g)))
            self._cache = items
        return self._cache

    def __getitem__(self, idx):
        base = self._pool()[idx % self.pool]
        item = dict(base)                                   # tensors are shared, the collate stacks (copies) them
        item["point_rgbs"] = base["ref_rgb"][None].expand(self.frames, -1, -1)
        item["obj_name"] = f"synthetic_{idx:06d}"
        return item
