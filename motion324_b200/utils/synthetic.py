"""Synthetic weights and inputs for benches, profiles and tests (SURVEY.md 8(d) "Synthetic inputs"): there is no network for
checkpoints or datasets, so every measurement runs on seeded random-init weights in the reference's ``state_dict`` layout and
on seeded random clips in the reference's sample-dict schema (dataset/dyscene.py:315-327).

``init_state_dict`` follows the reference's init statistics (Pcd_motion.py:283-342, transformer.py:15-25) with the key names and
shapes of SURVEY.md A.1 (incl. the hub DINOv2 ViT-B/14 names); the UNMODIFIED reference class loads it with ``strict=True``
(tests/golden/make_golden.py).  Plain data generation: no model arithmetic lives here.
"""
import math

import numpy as np
import torch

DEFAULT_CFG = dict(
    d=768, d_head=64, tokens=64, pcd_layers=4, n_layer=16, image_size=224, patch_size=14,
    frames=12, coord_mse_loss_weight=1.0,
)

DINO_EMBED_DIM, DINO_TRAIN_GRID, DINO_DEPTH = 768, 37, 12   # hub dinov2_vitb14: width, 518 / 14 position grid, blocks


def generate_pos_embed(T, H, W, embed_dim):
    """generate_pos_embed (Pcd_motion.py:230-266)."""
    def axis(n):
        if n > 1:
            return 2 * (torch.arange(n, dtype=torch.float32) / (n - 1)) - 1
        return torch.tensor([0.0], dtype=torch.float32)
    t, h, w = torch.meshgrid(axis(T), axis(H), axis(W), indexing="ij")
    pos = torch.stack([t, h, w], dim=-1)
    freq = 2.0 ** torch.linspace(0.0, 7.0, embed_dim // 6)
    pos = pos.unsqueeze(-1) * freq.view(1, 1, 1, 1, -1)
    pos = torch.cat([torch.sin(pos), torch.cos(pos)], dim=-1)
    return pos.reshape(1, -1, embed_dim)



def state_dict_spec(cfg=None):
    """(key, shape, kind) for every entry of the reference state_dict.  kind selects
    the initialiser: the reference's own init scheme (Pcd_motion.py:283-342, transformer.py:15-25)."""
    cfg = dict(DEFAULT_CFG, **(cfg or {}))
    d, dh, T = cfg["d"], cfg["d_head"], cfg["frames"]
    hp = cfg["image_size"] // cfg["patch_size"]
    spec = [("learnable_tokens", (1, cfg["tokens"], d), "randn"),
            ("special_token_0", (1, 4, d), "randn"), ("special_token_rest", (1, 4, d), "randn"),
            ("pos_embed", (1, T * hp * hp, d), "pos_embed"), ("point_embed.basis", (3, 24), "basis"),
            ("point_embed.mlp.weight", (d, 51), "kaiming"), ("point_embed.mlp.bias", (d,), "smallu"),
            ("point_normal_rgb_proj.weight", (d, d + 6), "n02"), ("point_normal_rgb_proj.bias", (d,), "zeros")]

    def cross(p, kind):
        s = [(p + "norm_q.weight", (d,), "ones"), (p + "norm_kv.weight", (d,), "ones"), (p + "norm2.weight", (d,), "ones")]
        for n in ("to_q", "to_k", "to_v", "fc"):
            s.append((p + f"attn.{n}.weight", (d, d), kind))
        s += [(p + "attn.q_norm.weight", (dh,), "ones"), (p + "attn.k_norm.weight", (dh,), "ones"),
              (p + "mlp.mlp.0.weight", (4 * d, d), kind), (p + "mlp.mlp.2.weight", (d, 4 * d), kind)]
        return s

    def selfb(p):
        return [(p + "norm1.weight", (d,), "ones"), (p + "attn.to_qkv.weight", (3 * d, d), "n02"),
                (p + "attn.fc.weight", (d, d), "n02"), (p + "attn.q_norm.weight", (dh,), "ones"),
                (p + "attn.k_norm.weight", (dh,), "ones"), (p + "norm2.weight", (d,), "ones"),
                (p + "mlp.mlp.0.weight", (4 * d, d), "n02"), (p + "mlp.mlp.2.weight", (d, 4 * d), "n02")]

    spec += cross("encoder_cross_attn.", "kaiming")
    for i in range(cfg["pcd_layers"]):
        spec += selfb(f"points_transformer_blocks.{i}.")
    for i in range(cfg["n_layer"] // 2):
        spec += selfb(f"global_transformer_blocks.{i}.")
    for i in range(cfg["n_layer"] // 2):
        spec += selfb(f"local_transformer_blocks.{i}.")
    spec += [("transformer_input_layernorm.weight", (d,), "ones")]
    spec += cross("decoder_cross_attn.", "kaiming")
    spec += [("shared_mlp_output.0.weight", (d,), "ones"), ("shared_mlp_output.0.bias", (d,), "zeros"),
             ("shared_mlp_output.1.weight", (d, d), "n02"), ("shared_mlp_output.1.bias", (d,), "zeros"),
             ("shared_mlp_output.3.weight", (3, d), "n02"), ("shared_mlp_output.3.bias", (3,), "zeros")]
    # DINOv2 ViT-B/14 (hub key names).  Random stand-in weights: N(0, 0.02) matrices, non-trivial
    # norms / biases / LayerScale so that every term of the arithmetic is exercised.
    p = "image_encoder.model."
    D = DINO_EMBED_DIM
    spec += [(p + "cls_token", (1, 1, D), "n02"), (p + "pos_embed", (1, 1 + DINO_TRAIN_GRID ** 2, D), "n02"),
             (p + "mask_token", (1, D), "zeros"),
             (p + "patch_embed.proj.weight", (D, 3, 14, 14), "n02"), (p + "patch_embed.proj.bias", (D,), "smallu")]
    for i in range(DINO_DEPTH):
        b = f"{p}blocks.{i}."
        spec += [(b + "norm1.weight", (D,), "near1"), (b + "norm1.bias", (D,), "smallu"),
                 (b + "attn.qkv.weight", (3 * D, D), "n02"), (b + "attn.qkv.bias", (3 * D,), "smallu"),
                 (b + "attn.proj.weight", (D, D), "n02"), (b + "attn.proj.bias", (D,), "smallu"),
                 (b + "ls1.gamma", (D,), "near1"),
                 (b + "norm2.weight", (D,), "near1"), (b + "norm2.bias", (D,), "smallu"),
                 (b + "mlp.fc1.weight", (4 * D, D), "n02"), (b + "mlp.fc1.bias", (4 * D,), "smallu"),
                 (b + "mlp.fc2.weight", (D, 4 * D), "n02"), (b + "mlp.fc2.bias", (D,), "smallu"),
                 (b + "ls2.gamma", (D,), "near1")]
    spec += [(p + "norm.weight", (D,), "near1"), (p + "norm.bias", (D,), "smallu")]
    return spec


def point_embed_basis():
    """PointEmbed basis (Pcd_motion.py:164-173): block-diagonal 2^k * pi, k=0..7, shape [3, 24]."""
    e = torch.pow(2, torch.arange(8)).float() * np.pi
    z = torch.zeros(8)
    return torch.stack([torch.cat([e, z, z]), torch.cat([z, e, z]), torch.cat([z, z, e])])


def init_state_dict(seed=0, cfg=None):
    """Deterministic random-init weights with the reference's key layout and init statistics.
    (Not bit-identical to constructing the reference class under torch.manual_seed: the reference
    class cannot be constructed on the GPU box; instead the reference LOADS this dict, strict.)"""
    cfg = dict(DEFAULT_CFG, **(cfg or {}))
    g = torch.Generator().manual_seed(seed)
    hp = cfg["image_size"] // cfg["patch_size"]
    sd = {}
    for key, shape, kind in state_dict_spec(cfg):
        if kind == "randn":
            t = torch.randn(shape, generator=g)
        elif kind == "n02":
            t = torch.randn(shape, generator=g) * 0.02
        elif kind == "kaiming":  # nn.Linear default: U(-1/sqrt(fan_in), 1/sqrt(fan_in))
            bound = 1.0 / math.sqrt(shape[-1])
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif kind == "smallu":
            t = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
        elif kind == "near1":
            t = 1.0 + (torch.rand(shape, generator=g) * 2 - 1) * 0.1
        elif kind == "ones":
            t = torch.ones(shape)
        elif kind == "zeros":
            t = torch.zeros(shape)
        elif kind == "pos_embed":
            t = generate_pos_embed(cfg["frames"], hp, hp, cfg["d"])
        elif kind == "basis":
            t = point_embed_basis()
        else:
            raise ValueError(kind)
        sd[key] = t.float()
    return sd


def make_inputs(seed=1, B=1, T=1, N=512, S=512, H=224, W=224, with_gt=True):
    """Synthetic sample dict (SURVEY.md 8(d) 'Synthetic inputs'; schema dataset/dyscene.py:315-327)."""
    g = torch.Generator().manual_seed(seed)
    u = lambda *s: torch.rand(*s, generator=g)
    n = lambda *s: torch.randn(*s, generator=g)
    unit = lambda x: x / x.norm(dim=-1, keepdim=True).clamp_min(1e-8)
    sample = {
        "ref_shape_pcd": u(B, S, 3) - 0.5, "ref_shape_normals": unit(n(B, S, 3)), "ref_shape_rgbs": u(B, S, 3),
        "ref_pcd": u(B, N, 3) - 0.5, "ref_normal": unit(n(B, N, 3)), "ref_rgb": u(B, N, 3),
        "rgb_video": u(B, T, H, W, 3),
    }
    if with_gt:
        sample["point_clouds"] = sample["ref_pcd"][:, None] + 0.05 * n(B, T, N, 3)
    return sample
