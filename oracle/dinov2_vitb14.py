"""ORACLE (test infrastructure, never imported by the product path).

CPU restatement of the third-party image encoder the reference pulls at run time:
``torch.hub.load('facebookresearch/dinov2', 'dinov2_vitb14')`` (reference call site
``model/image_encoder/dinov2.py:44``; un-vendored, unpinned hub default branch).  The
reference only touches three things on the returned module: ``.patch_size``
(``dinov2.py:46-50``), ``.embed_dim`` (``dinov2.py:117-119``) and
``.forward_features(x)['x_norm_patchtokens']`` (``dinov2.py:99-103``).

The arithmetic below restates upstream ``dinov2/models/vision_transformer.py`` +
``dinov2/layers/{patch_embed,block,attention,mlp,layer_scale}.py`` for the ViT-B/14 hub
configuration: patch 14, embed 768, depth 12, heads 12, mlp ratio 4, qkv bias, proj bias,
ffn bias, LayerNorm eps 1e-6, LayerScale (init 1.0), no register tokens, position table
trained at 518x518 (1 + 37*37 = 1370 entries), bicubic position interpolation with
``interpolate_offset = 0.1`` and ``interpolate_antialias = False``.  Parameter names follow the
hub checkpoint (``cls_token``, ``pos_embed``, ``mask_token``, ``patch_embed.proj.*``,
``blocks.N.{norm1,attn.qkv,attn.proj,ls1.gamma,norm2,mlp.fc1,mlp.fc2,ls2.gamma}``, ``norm``)
so that a state_dict saved by the reference (keys ``image_encoder.model.*``) loads unchanged.

PARITY UNPINNED at this boundary: nothing inside /root/reference pins DINOv2 numerics (the
upstream source and weights are absent and there is no network).  The restatement is
cross-checked against ``transformers.models.dinov2`` (tests/test_oracle_dino.py), which is an
independent implementation of the same network.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

EMBED_DIM = 768
DEPTH = 12
NUM_HEADS = 12
PATCH = 14
TRAIN_GRID = 37  # 518 / 14
LN_EPS = 1e-6
INTERPOLATE_OFFSET = 0.1
FUSED_SDPA = False   # reference_gpu leg only: upstream's MemEffAttention runs a fused attention kernel on CUDA


class _Attention(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.num_heads = heads
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim, bias=True)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads)
        q, k, v = qkv.unbind(2)  # [B, N, H, Dh]
        q, k, v = (t.transpose(1, 2) for t in (q, k, v))
        if FUSED_SDPA:
            return self.proj(F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, N, C))
        scale = (C // self.num_heads) ** -0.5
        attn = (q * scale) @ k.transpose(-2, -1)
        attn = attn.softmax(dim=-1)
        x = (attn @ v).transpose(1, 2).reshape(B, N, C)
        return self.proj(x)


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden, bias=True)
        self.fc2 = nn.Linear(hidden, dim, bias=True)

    def forward(self, x):
        return self.fc2(F.gelu(self.fc1(x)))


class _LayerScale(nn.Module):
    def __init__(self, dim, init=1.0):
        super().__init__()
        self.gamma = nn.Parameter(init * torch.ones(dim))

    def forward(self, x):
        return x * self.gamma


class _Block(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=LN_EPS)
        self.attn = _Attention(dim, heads)
        self.ls1 = _LayerScale(dim)
        self.norm2 = nn.LayerNorm(dim, eps=LN_EPS)
        self.mlp = _Mlp(dim, dim * 4)
        self.ls2 = _LayerScale(dim)

    def forward(self, x):
        x = x + self.ls1(self.attn(self.norm1(x)))
        x = x + self.ls2(self.mlp(self.norm2(x)))
        return x


class _PatchEmbed(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.proj = nn.Conv2d(3, dim, kernel_size=PATCH, stride=PATCH)

    def forward(self, x):
        x = self.proj(x)  # B, C, h, w
        return x.flatten(2).transpose(1, 2)


def interpolate_pos_embed(pos_embed, npatch_h, npatch_w, offset=INTERPOLATE_OFFSET):
    """Upstream ``DinoVisionTransformer.interpolate_pos_encoding`` (offset variant, the hub default
    ``interpolate_offset = 0.1``): bicubic resize of the 37x37 patch-position table with scale factor
    (n + 0.1) / 37, class position passed through.  Returns [1, 1 + h*w, C] in fp32.
    ``offset=0`` is upstream's other branch (``interpolate_offset = 0`` -> ``size=(h, w)``), which is also what
    ``transformers`` implements; tests/test_oracle_dino.py pins both branches."""
    pos_embed = pos_embed.float()
    N = pos_embed.shape[1] - 1
    M = int(math.sqrt(N))
    dim = pos_embed.shape[-1]
    if npatch_h * npatch_w == N and npatch_h == npatch_w:
        return pos_embed
    class_pos = pos_embed[:, :1]
    patch_pos = pos_embed[:, 1:].reshape(1, M, M, dim).permute(0, 3, 1, 2)
    if offset:
        sx = float(npatch_w + offset) / M
        sy = float(npatch_h + offset) / M
        patch_pos = F.interpolate(patch_pos, scale_factor=(sy, sx), mode="bicubic", antialias=False)
    else:
        patch_pos = F.interpolate(patch_pos, size=(npatch_h, npatch_w), mode="bicubic", antialias=False)
    assert patch_pos.shape[-2:] == (npatch_h, npatch_w)
    patch_pos = patch_pos.permute(0, 2, 3, 1).reshape(1, -1, dim)
    return torch.cat([class_pos, patch_pos], dim=1)


class DinoV2ViTB14(nn.Module):
    """Stand-in for the hub module; same attribute surface the reference uses."""

    patch_size = PATCH
    embed_dim = EMBED_DIM

    def __init__(self):
        super().__init__()
        self.cls_token = nn.Parameter(torch.zeros(1, 1, EMBED_DIM))
        self.pos_embed = nn.Parameter(torch.zeros(1, 1 + TRAIN_GRID * TRAIN_GRID, EMBED_DIM))
        self.mask_token = nn.Parameter(torch.zeros(1, EMBED_DIM))
        self.patch_embed = _PatchEmbed(EMBED_DIM)
        self.blocks = nn.ModuleList([_Block(EMBED_DIM, NUM_HEADS) for _ in range(DEPTH)])
        self.norm = nn.LayerNorm(EMBED_DIM, eps=LN_EPS)

    def prepare_tokens(self, x):
        B, _, H, W = x.shape
        t = self.patch_embed(x)
        t = torch.cat([self.cls_token.expand(B, -1, -1), t], dim=1)
        return t + interpolate_pos_embed(self.pos_embed, H // PATCH, W // PATCH).to(t.dtype)

    def forward_features(self, x):
        t = self.prepare_tokens(x)
        for blk in self.blocks:
            t = blk(t)
        t = self.norm(t)
        return {"x_norm_clstoken": t[:, 0], "x_norm_patchtokens": t[:, 1:], "x_prenorm": None}

    def forward(self, x):
        return self.forward_features(x)["x_norm_clstoken"]
