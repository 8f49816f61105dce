"""ORACLE (test infrastructure; only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs import this; the product path, bench.py's own arm and scripts/ never do --
the seeded weights / inputs generator they need lives in motion324_b200/utils/synthetic.py and is re-exported here).

CPU restatement, in plain functional torch on a flat ``state_dict``, of the reference hot
path ``Motion_Latent_Model.forward`` (/root/reference/model/Pcd_motion.py:450-598) and the
operators under it (/root/reference/model/transformer.py, model/loss.py,
model/image_encoder/dinov2.py).  Every function cites the reference lines it follows.

Pinning: the reference ships no tests / golden vectors, so the oracle is
pinned against OUTPUTS OF THE REFERENCE ITSELF, produced in the build container by
tests/golden/make_golden.py (which imports /root/reference with the three shims of
oracle/ref_shims.py) and committed under tests/golden/*.npz; tests/test_oracle_golden.py
replays them.  The DINOv2 arithmetic lives in an un-vendored third party (see
oracle/dinov2_vitb14.py): parity at that boundary is UNPINNED by the reference.

dtype: everything runs in the dtype of the tensors passed in (fp32 = the parity target,
fp64 = error budgeting).  ``prec`` optionally emulates the product's operand rounding
(fp16 tensor-core inputs, fp32 accumulate) for error budgeting on CPU.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import dinov2_vitb14 as dino

# ----------------------------------------------------------------------------- config

from motion324_b200.utils.synthetic import DEFAULT_CFG  # noqa: E402


class Prec:
    """Operand-rounding model. ``mode``: None (exact), 'fp16', 'bf16', 'tf32'."""

    def __init__(self, mode=None):
        self.mode = mode

    def r(self, x):
        if self.mode is None:
            return x
        if self.mode == "fp16":
            return x.to(torch.float16).to(x.dtype)
        if self.mode == "bf16":
            return x.to(torch.bfloat16).to(x.dtype)
        if self.mode == "tf32":
            xi = x.float().contiguous().view(torch.int32)
            xi = (xi + 0x1000) & ~0x1FFF
            return xi.view(torch.float32).to(x.dtype)
        if self.mode == "fp16x2":  # two-term split: hi + lo, each fp16 (≈21-bit mantissa)
            hi = x.to(torch.float16).to(x.dtype)
            lo = (x - hi).to(torch.float16).to(x.dtype)
            return hi + lo
        raise ValueError(self.mode)


EXACT = Prec(None)


def product_prec():
    """The product's operand-precision plan (DESIGN.md 'Precision'): fp16 tensor-core operands with
    fp32 accumulation everywhere, except the three GEMMs that dominate the output error at
    negligible FLOP cost, which run as 3-pass split-fp16 (hi/lo) GEMMs, and the 768->3 head,
    which is an fp32 CUDA-core dot product."""
    h = Prec("fp16")
    return dict(shape=h, dino=h, trunk=h, dec=h, pf=Prec("fp16x2"), head1=Prec("fp16x2"), head2=EXACT)

# ----------------------------------------------------------------------------- operators


def linear(x, w, b=None, prec=EXACT):
    """nn.Linear: x @ w.T + b (operands optionally rounded, accumulate in x.dtype)."""
    y = prec.r(x) @ prec.r(w).t()
    return y if b is None else y + b


def layer_norm(x, w, b=None, eps=1e-5):
    """nn.LayerNorm over the last dim (transformer.py:345-357,400,411: bias=False, eps 1e-5)."""
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def rms_norm(x, w, eps=1e-5):
    """RMSNorm (transformer.py:30-42): x * rsqrt(mean(x^2) + eps) * w, over the head dim."""
    return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * w


def attention(q, k, v, prec=EXACT):
    """xops.memory_efficient_attention(q, k, v, attn_bias=None, p=0) with layout [B, L, H, Dh]
    (transformer.py:134-139, 209-214): softmax(q k^T / sqrt(Dh)) v, no mask."""
    B, Lq, H, Dh = q.shape
    q_, k_, v_ = (prec.r(t).transpose(1, 2) for t in (q, k, v))  # [B,H,L,Dh]
    s = (q_ @ k_.transpose(-2, -1)) * (Dh ** -0.5)
    p = torch.softmax(s, dim=-1)
    if prec.mode is not None:
        # the product normalises after P.V: unnormalised exp rounded to fp16, fp32 row sum
        m = s.max(dim=-1, keepdim=True).values
        e = torch.exp(s - m)
        o = (prec.r(e) @ v_) / e.sum(-1, keepdim=True)
    else:
        o = p @ v_
    return o.transpose(1, 2)  # [B, Lq, H, Dh]


def gelu(x):
    return F.gelu(x)  # exact erf form (nn.GELU default, transformer.py:58)


def mlp(x, sd, pfx, prec=EXACT):
    """MLP (transformer.py:46-81): Linear(d,4d,no bias) -> GELU -> Linear(4d,d,no bias)."""
    return linear(gelu(linear(x, sd[pfx + "mlp.0.weight"], None, prec)), sd[pfx + "mlp.2.weight"], None, prec)


def self_attention_block(x, sd, pfx, d_head, prec=EXACT):
    """QK_Norm_TransformerBlock.forward (transformer.py:420-423) with QK_Norm_SelfAttention
    (transformer.py:191-219): x += fc(MEA(rms(q), rms(k), v)); x += mlp(LN(x))."""
    B, L, C = x.shape
    H = C // d_head
    h = layer_norm(x, sd[pfx + "norm1.weight"])
    qkv = linear(h, sd[pfx + "attn.to_qkv.weight"], None, prec)
    q, k, v = qkv.chunk(3, dim=-1)
    q, k, v = (t.reshape(B, L, H, d_head) for t in (q, k, v))
    q = rms_norm(q, sd[pfx + "attn.q_norm.weight"])
    k = rms_norm(k, sd[pfx + "attn.k_norm.weight"])
    a = attention(q, k, v, prec).reshape(B, L, C)
    x = x + linear(a, sd[pfx + "attn.fc.weight"], None, prec)
    x = x + mlp(layer_norm(x, sd[pfx + "norm2.weight"]), sd, pfx + "mlp.", prec)
    return x


def cross_attention_block(query, kv, sd, pfx, d_head, prec=EXACT):
    """QK_Norm_CrossAttentionBlock.forward (transformer.py:365-377) with QK_Norm_CrossAttention
    (transformer.py:123-144); key is value at both call sites (Pcd_motion.py:462, 556-560)."""
    B, Lq, C = query.shape
    Lk = kv.shape[1]
    H = C // d_head
    qn = layer_norm(query, sd[pfx + "norm_q.weight"])
    kn = layer_norm(kv, sd[pfx + "norm_kv.weight"])
    q = linear(qn, sd[pfx + "attn.to_q.weight"], None, prec).reshape(B, Lq, H, d_head)
    k = linear(kn, sd[pfx + "attn.to_k.weight"], None, prec).reshape(B, Lk, H, d_head)
    v = linear(kn, sd[pfx + "attn.to_v.weight"], None, prec).reshape(B, Lk, H, d_head)
    q = rms_norm(q, sd[pfx + "attn.q_norm.weight"])
    k = rms_norm(k, sd[pfx + "attn.k_norm.weight"])
    a = attention(q, k, v, prec).reshape(B, Lq, C)
    x = query + linear(a, sd[pfx + "attn.fc.weight"], None, prec)
    x = x + mlp(layer_norm(x, sd[pfx + "norm2.weight"]), sd, pfx + "mlp.", prec)
    return x


def point_embed(x, sd, prec=EXACT):
    """PointEmbed.forward (Pcd_motion.py:177-187): Linear(51,768)(cat[sin(x@basis), cos(x@basis), x])."""
    proj = torch.einsum("bnd,de->bne", x, sd["point_embed.basis"].to(x.dtype))
    emb = torch.cat([proj.sin(), proj.cos(), x], dim=2)
    return linear(emb, sd["point_embed.mlp.weight"], sd["point_embed.mlp.bias"], prec)


def point_features(xyz, normal, rgb, sd, prec=EXACT):
    """Pcd_motion.py:456-459 and :550-553: point_normal_rgb_proj(cat[point_embed(xyz), normal, rgb])."""
    e = point_embed(xyz, sd, prec)
    return linear(torch.cat([e, normal, rgb], dim=-1), sd["point_normal_rgb_proj.weight"],
                  sd["point_normal_rgb_proj.bias"], prec)


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


def resize_pos_embed(posemb, src_shape, target_shape):
    """resize_pos_embed (Pcd_motion.py:221-228): trilinear, align_corners=False."""
    p = posemb.reshape(1, src_shape[0], src_shape[1], src_shape[2], -1).permute(0, 4, 1, 2, 3)
    p = F.interpolate(p, size=target_shape, mode="trilinear", align_corners=False)
    return p.permute(0, 2, 3, 4, 1).reshape(1, target_shape[0] * target_shape[1] * target_shape[2], -1)


_MEAN = [0.485, 0.456, 0.406]
_STD = [0.229, 0.224, 0.225]


def dino_forward(images, sd, pfx="image_encoder.model.", prec=EXACT):
    """DinoEncoder.forward (image_encoder/dinov2.py:65-124) + hub ViT-B/14 forward_features
    (restated in oracle/dinov2_vitb14.py).  images [B,3,224,224] in [0,1] -> [B,256,768]."""
    dt = images.dtype
    mean = torch.tensor(_MEAN, dtype=dt, device=images.device).view(1, 3, 1, 1)
    std = torch.tensor(_STD, dtype=dt, device=images.device).view(1, 3, 1, 1)
    x = (images - mean) / std
    B, _, H, W = x.shape
    P = dino.PATCH
    hp, wp = H // P, W // P
    # patch embed (Conv2d k=14 s=14) as im2col GEMM
    patches = x.reshape(B, 3, hp, P, wp, P).permute(0, 2, 4, 1, 3, 5).reshape(B, hp * wp, 3 * P * P)
    wpe = sd[pfx + "patch_embed.proj.weight"].reshape(dino.EMBED_DIM, -1)
    t = linear(patches, wpe, sd[pfx + "patch_embed.proj.bias"], prec)
    t = torch.cat([sd[pfx + "cls_token"].to(dt).expand(B, -1, -1), t], dim=1)
    t = t + dino.interpolate_pos_embed(sd[pfx + "pos_embed"], hp, wp).to(dt)
    Hh = dino.NUM_HEADS
    Dh = dino.EMBED_DIM // Hh
    for i in range(dino.DEPTH):
        b = f"{pfx}blocks.{i}."
        h = layer_norm(t, sd[b + "norm1.weight"], sd[b + "norm1.bias"], dino.LN_EPS)
        qkv = linear(h, sd[b + "attn.qkv.weight"], sd[b + "attn.qkv.bias"], prec)
        L = qkv.shape[1]
        q, k, v = qkv.reshape(B, L, 3, Hh, Dh).unbind(2)
        a = attention(q, k, v, prec).reshape(B, L, Hh * Dh)
        t = t + sd[b + "ls1.gamma"] * linear(a, sd[b + "attn.proj.weight"], sd[b + "attn.proj.bias"], prec)
        h = layer_norm(t, sd[b + "norm2.weight"], sd[b + "norm2.bias"], dino.LN_EPS)
        h = linear(gelu(linear(h, sd[b + "mlp.fc1.weight"], sd[b + "mlp.fc1.bias"], prec)),
                   sd[b + "mlp.fc2.weight"], sd[b + "mlp.fc2.bias"], prec)
        t = t + sd[b + "ls2.gamma"] * h
    t = layer_norm(t, sd[pfx + "norm.weight"], sd[pfx + "norm.bias"], dino.LN_EPS)
    return t[:, 1:]


def mse_loss(pred, target, weight):
    """MSELossComputer.forward (loss.py:24-66)."""
    if not (pred.ndim == 4 and target.ndim == 4 and pred.shape == target.shape):
        raise ValueError("Shape mismatch or invalid shape for coordinate MSE. Expected both tensors of shape "
                         f"(B, T, N, C). Got pred: {pred.shape}, target: {target.shape}")
    mse = ((pred - target) ** 2).mean() if weight > 0.0 else torch.zeros((), dtype=pred.dtype, device=pred.device)
    return dict(coord_mse_loss=mse, loss=weight * mse)


# ----------------------------------------------------------------------------- the hot path


def forward(sd, sample, cfg=None, training=False, prec=EXACT, return_stages=False):
    """Motion_Latent_Model.forward (Pcd_motion.py:450-598), eval semantics (pos_drop off).

    sd: flat state_dict (reference key layout); sample: reference sample dict
    (SURVEY.md 8(a) a14).  Returns dict(pcd_moved[B,T,N,3], loss_metrics?) (+ stages)."""
    cfg = dict(DEFAULT_CFG, **(cfg or {}))
    d, dh, ntok = cfg["d"], cfg["d_head"], cfg["tokens"]
    # prec may be one Prec or a per-stage dict {'shape','dino','trunk','dec'} (error budgeting)
    if not isinstance(prec, dict):
        prec = dict(shape=prec, dino=prec, trunk=prec, dec=prec)
    p_shape, p_dino, p_trunk, p_dec = (prec.get(k, EXACT) for k in ("shape", "dino", "trunk", "dec"))
    p_pf = prec.get("pf", p_dec)      # point_embed.mlp + point_normal_rgb_proj (both call sites)
    p_h1 = prec.get("head1", p_dec)   # shared_mlp_output.1
    p_h2 = prec.get("head2", p_dec)   # shared_mlp_output.3
    dt = sample["ref_pcd"].dtype
    sd = {k: (v.to(dt) if v.is_floating_point() else v) for k, v in sd.items()}
    stages = {}
    B, N_points = sample["ref_pcd"].shape[:2]

    # A. shape encoder (Pcd_motion.py:456-464)
    shape_feat = point_features(sample["ref_shape_pcd"], sample["ref_shape_normals"], sample["ref_shape_rgbs"], sd, prec.get("pf", p_shape))
    query_tokens = sd["learnable_tokens"].expand(B, -1, -1)
    mesh_feat = cross_attention_block(query_tokens, shape_feat, sd, "encoder_cross_attn.", dh, p_shape)
    for i in range(cfg["pcd_layers"]):
        mesh_feat = self_attention_block(mesh_feat, sd, f"points_transformer_blocks.{i}.", dh, p_shape)
    stages["mesh_feat"] = mesh_feat

    # B. video encoder (Pcd_motion.py:466-490)
    rgb = sample["rgb_video"]
    Bv, T, Hin, Win, _ = rgb.shape
    x = rgb.permute(0, 1, 4, 2, 3).reshape(Bv * T, 3, Hin, Win)
    S = cfg["image_size"]
    x = F.interpolate(x, (S, S), mode="bilinear", align_corners=False)
    img_tok = dino_forward(x, sd, prec=p_dino)  # [B*T, 256, 768]
    stages["dino_tokens"] = img_tok
    hp = S // cfg["patch_size"]
    x = img_tok.reshape(Bv, T, hp, hp, d).permute(0, 4, 1, 2, 3).flatten(2).transpose(1, 2)  # [B, T*256, d]
    lat_T = cfg["frames"]
    pos = sd["pos_embed"]
    if T != lat_T:
        pos = resize_pos_embed(pos, (lat_T, hp, hp), (T, hp, hp))
    x = x + pos
    video_tokens = x.reshape(B, T, hp * hp, d)

    # C. token assembly (Pcd_motion.py:495-509)
    sp0 = sd["special_token_0"].expand(B, 4, -1)
    spr = sd["special_token_rest"].expand(B, 4, -1)
    specials = torch.stack([sp0] + [spr] * (T - 1), dim=1)
    tokens = torch.cat([specials, mesh_feat.unsqueeze(1).expand(B, T, ntok, d), video_tokens], dim=2)
    tokens = layer_norm(tokens, sd["transformer_input_layernorm.weight"])
    L = tokens.shape[2]
    stages["trunk_in"] = tokens

    # D. trunk (pass_alternating_attention, Pcd_motion.py:394-409)
    for i in range(cfg["n_layer"] // 2):
        tokens = self_attention_block(tokens.reshape(B, T * L, d), sd, f"global_transformer_blocks.{i}.", dh, p_trunk)
        tokens = self_attention_block(tokens.reshape(B * T, L, d), sd, f"local_transformer_blocks.{i}.", dh, p_trunk)
        tokens = tokens.reshape(B, T, L, d)
    stages["trunk_out"] = tokens
    pcd_tokens = tokens[:, :, 4:4 + ntok, :]  # Pcd_motion.py:520

    # E. decoder (decode_chunk, Pcd_motion.py:529-564; chunks of 4096 in eval, :566-575)
    def decode_chunk(pcd, normal, rgbs):
        feat = point_features(pcd, normal, rgbs, sd, p_pf)  # identical for every t (:534-553)
        outs = []
        for t in range(T):
            dec = cross_attention_block(feat, pcd_tokens[:, t], sd, "decoder_cross_attn.", dh, p_dec)
            h = layer_norm(dec, sd["shared_mlp_output.0.weight"], sd["shared_mlp_output.0.bias"])
            h = gelu(linear(h, sd["shared_mlp_output.1.weight"], sd["shared_mlp_output.1.bias"], p_h1))
            outs.append(linear(h, sd["shared_mlp_output.3.weight"], sd["shared_mlp_output.3.bias"], p_h2))
        return torch.stack(outs, dim=1)  # [B, T, n, 3]

    chunk = 4096
    if (not training) and N_points > chunk:
        parts = [decode_chunk(sample["ref_pcd"][:, i:i + chunk], sample["ref_normal"][:, i:i + chunk],
                              sample["ref_rgb"][:, i:i + chunk]) for i in range(0, N_points, chunk)]
        out = torch.cat(parts, dim=2)
    else:
        out = decode_chunk(sample["ref_pcd"], sample["ref_normal"], sample["ref_rgb"])

    result = dict(pcd_moved=out)
    if "point_clouds" in sample:  # Pcd_motion.py:582-592
        lm = mse_loss(out, sample["point_clouds"], cfg["coord_mse_loss_weight"])
        result["loss_metrics"] = dict(loss=lm["loss"], xyz_loss=lm["coord_mse_loss"])
    if return_stages:
        result["stages"] = stages
    return result


# ----------------------------------------------------------------------------- weights / inputs


from motion324_b200.utils.synthetic import (  # noqa: E402,F401  seeded weights / inputs generator (plain data, lives in the package)
    state_dict_spec, point_embed_basis, init_state_dict, make_inputs)


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
