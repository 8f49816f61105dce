"""Drop-in for the reference plugin ``model.Pcd_motion.Motion_Latent_Model``
(/root/reference/model/Pcd_motion.py:268-598): same constructor ``Cls(config)``, same ``state_dict`` key layout
(proven by tests/golden: the reference class loads our dict with strict=True), same
``forward(sample) -> EasyDict{input_data, pcd_moved[B,T,N,3], loss_metrics{loss, xyz_loss}}``.

Every arithmetic step of the forward runs in libm324.so (hand-written sm_100a kernels, C ABI in include/m324.h);
this file only owns parameters, workspaces and the launch order.  There is no PyTorch / CPU fallback: a missing
library or a non-CUDA input raises.  Select it from the reference's entry points with
``model.class_name=motion324_b200.model.Pcd_motion.Motion_Latent_Model`` (train.py:84-86,
scripts/inference_with_video_mesh.py:309-311).

Inference (``eval()`` or ``torch.no_grad()``): forward + loss.  Training (``train()`` with grad enabled, train.py:143-170):
the forward keeps its activations and the hand-written backward (model/train_path.py) runs in the same
call; ``ret.loss_metrics.loss`` is a differentiable scalar whose ``backward()`` hands every trainable parameter its
gradient (so ``DDP(model)`` and ``loss / grad_accum_steps`` work unchanged), and ``forward_backward(sample)`` is the
direct entry that leaves the gradients in one flat buffer for a single NCCL all-reduce.
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..utils.easydict import EasyDict as edict
from .train_path import TrainPath

def frame_shard(T, rank, world):
    """Frames [first, first + count) of a T-frame clip owned by `rank` under frame_parallel(): equal contiguous blocks, so that
    an all-gather in rank order restores frame order.  T must divide by the group size."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank {rank} of {world}")
    if T % world != 0:
        raise ValueError(f"frame_parallel needs a frame count that divides by the group size; got T={T}, world={world}")
    return rank * (T // world), T // world


DINO_DEPTH, DINO_DIM, DINO_GRID, DINO_EPS = 12, 768, 37, 1e-6
KP_EMB, KP_FEAT, KP_PATCH = 64, 832, 640   # K paddings (multiples of 64) of the 51-, 774- and 588-wide operands


def _cfg_get(node, key, default=None):
    if isinstance(node, dict):
        return node.get(key, default)
    return getattr(node, key, default)


# --------------------------------------------------------------------------------------------- parameter containers
# Plain nn containers reproduce the reference's state_dict keys; their forward() is never used.


class _RMSNormP(nn.Module):  # transformer.py:30-42
    def __init__(self, dim):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))


class _MLPP(nn.Module):  # transformer.py:46-81 -> keys mlp.mlp.0.weight / mlp.mlp.2.weight
    def __init__(self, dim):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(dim, 4 * dim, bias=False), nn.GELU(), nn.Linear(4 * dim, dim, bias=False),
                                 nn.Dropout(0.0))


class _SelfAttnP(nn.Module):  # transformer.py:146-189
    def __init__(self, dim, head_dim):
        super().__init__()
        self.to_qkv = nn.Linear(dim, 3 * dim, bias=False)
        self.fc = nn.Linear(dim, dim, bias=False)
        self.q_norm = _RMSNormP(head_dim)
        self.k_norm = _RMSNormP(head_dim)


class _CrossAttnP(nn.Module):  # transformer.py:84-121
    def __init__(self, dim, head_dim):
        super().__init__()
        self.to_q = nn.Linear(dim, dim, bias=False)
        self.to_k = nn.Linear(dim, dim, bias=False)
        self.to_v = nn.Linear(dim, dim, bias=False)
        self.fc = nn.Linear(dim, dim, bias=False)
        self.q_norm = _RMSNormP(head_dim)
        self.k_norm = _RMSNormP(head_dim)


class _SelfBlockP(nn.Module):  # transformer.py:379-417
    def __init__(self, dim, head_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, bias=False)
        self.attn = _SelfAttnP(dim, head_dim)
        self.norm2 = nn.LayerNorm(dim, bias=False)
        self.mlp = _MLPP(dim)


class _CrossBlockP(nn.Module):  # transformer.py:324-363
    def __init__(self, dim, head_dim):
        super().__init__()
        self.norm_q = nn.LayerNorm(dim, bias=False)
        self.norm_kv = nn.LayerNorm(dim, bias=False)
        self.attn = _CrossAttnP(dim, head_dim)
        self.norm2 = nn.LayerNorm(dim, bias=False)
        self.mlp = _MLPP(dim)


class _PointEmbedP(nn.Module):  # Pcd_motion.py:157-175
    def __init__(self, dim):
        super().__init__()
        e = torch.pow(2, torch.arange(8)).float() * np.pi
        z = torch.zeros(8)
        self.register_buffer("basis", torch.stack([torch.cat([e, z, z]), torch.cat([z, e, z]), torch.cat([z, z, e])]))
        self.mlp = nn.Linear(51, dim)


class _DinoAttnP(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.qkv = nn.Linear(dim, 3 * dim)
        self.proj = nn.Linear(dim, dim)


class _DinoMlpP(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.fc1 = nn.Linear(dim, 4 * dim)
        self.fc2 = nn.Linear(4 * dim, dim)


class _GammaP(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(dim))


class _DinoBlockP(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=DINO_EPS)
        self.attn = _DinoAttnP(dim)
        self.ls1 = _GammaP(dim)
        self.norm2 = nn.LayerNorm(dim, eps=DINO_EPS)
        self.mlp = _DinoMlpP(dim)
        self.ls2 = _GammaP(dim)


class _PatchEmbedP(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.proj = nn.Conv2d(3, dim, kernel_size=14, stride=14)


class _DinoViTP(nn.Module):
    """Parameter layout of hub ``dinov2_vitb14`` (image_encoder/dinov2.py:44); frozen."""
    patch_size = 14
    embed_dim = DINO_DIM

    def __init__(self):
        super().__init__()
        self.cls_token = nn.Parameter(torch.zeros(1, 1, DINO_DIM))
        self.pos_embed = nn.Parameter(torch.zeros(1, 1 + DINO_GRID * DINO_GRID, DINO_DIM))
        self.mask_token = nn.Parameter(torch.zeros(1, DINO_DIM))
        self.patch_embed = _PatchEmbedP(DINO_DIM)
        self.blocks = nn.ModuleList([_DinoBlockP(DINO_DIM) for _ in range(DINO_DEPTH)])
        self.norm = nn.LayerNorm(DINO_DIM, eps=DINO_EPS)


class _DinoEncoderP(nn.Module):  # image_encoder/dinov2.py:39-63,126-131
    def __init__(self):
        super().__init__()
        self.model = _DinoViTP()
        for p in self.model.parameters():
            p.requires_grad = False


def _init_weights(module, std=0.02):  # transformer.py:15-25
    if isinstance(module, (nn.Linear, nn.Embedding)):
        nn.init.normal_(module.weight, mean=0.0, std=std)
        if isinstance(module, nn.Linear) and module.bias is not None:
            nn.init.zeros_(module.bias)


def generate_pos_embed(T, H, W, embed_dim):  # Pcd_motion.py:230-266 (init-time buffer)
    def axis(n):
        return 2 * (torch.arange(n, dtype=torch.float32) / (n - 1)) - 1 if n > 1 else torch.tensor([0.0])
    t, h, w = torch.meshgrid(axis(T), axis(H), axis(W), indexing="ij")
    pos = torch.stack([t, h, w], dim=-1).unsqueeze(-1) * (2.0 ** torch.linspace(0.0, 7.0, embed_dim // 6)).view(1, 1, 1, 1, -1)
    return torch.cat([torch.sin(pos), torch.cos(pos)], dim=-1).reshape(1, -1, embed_dim)


# --------------------------------------------------------------------------------------------- the model


class Motion_Latent_Model(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        mcfg, tcfg = config.model, config.training
        if _cfg_get(tcfg, "coord_mse_loss_weight") is None:  # model/loss.py:18-22
            raise ValueError("Configuration must have 'config.training.coord_mse_loss_weight' defined.")
        self.feat_dim = mcfg.feat_dim
        vcfg = mcfg.video_encoder
        tr, tok = vcfg.transformer, vcfg.image_tokenizer
        d, dh = tr.d, tr.d_head
        if not (d == 768 and dh == 64 and self.feat_dim == 768 and _cfg_get(tr, "use_qk_norm", True)):
            raise ValueError("libm324 kernels are built for d=768, d_head=64, feat_dim=768, use_qk_norm=True")
        self.d, self.dh, self.H = d, dh, d // dh
        self.image_size = _cfg_get(tok, "image_size", 224)
        self.patch_size = _cfg_get(tok, "patch_size", 14)
        if self.image_size != 224 or self.patch_size != 14:
            raise ValueError("DinoEncoder is fixed to 224x224 / patch 14 (image_encoder/dinov2.py:55-58)")
        self.video_length = tcfg.frames
        self.hp = self.image_size // self.patch_size
        self.latent_length = self.video_length // _cfg_get(tok, "patch_length", 1)
        self.register_buffer("pos_embed", generate_pos_embed(self.latent_length, self.hp, self.hp, d))
        self.drop_rate = _cfg_get(tr, "drop_rate", 0.1)

        self.point_embed = _PointEmbedP(d)
        self.point_normal_rgb_proj = nn.Linear(d + 6, d)
        self.point_normal_rgb_proj.apply(_init_weights)
        self.num_learnable_tokens = mcfg.tokens
        self.learnable_tokens = nn.Parameter(torch.randn(1, self.num_learnable_tokens, d))
        self.special_token_0 = nn.Parameter(torch.randn(1, 4, d))
        self.special_token_rest = nn.Parameter(torch.randn(1, 4, d))
        self.encoder_cross_attn = _CrossBlockP(d, dh)
        self.points_transformer_blocks = nn.ModuleList([_SelfBlockP(d, dh) for _ in range(mcfg.pcd_layers)])
        self.points_transformer_blocks.apply(_init_weights)
        self.image_encoder = _DinoEncoderP()
        self.alternating_layers = _cfg_get(tr, "n_layer", 12)
        assert self.alternating_layers % 2 == 0, "Alternating layers should be even."
        self.global_transformer_blocks = nn.ModuleList([_SelfBlockP(d, dh) for _ in range(self.alternating_layers // 2)])
        self.global_transformer_blocks.apply(_init_weights)
        self.local_transformer_blocks = nn.ModuleList([_SelfBlockP(d, dh) for _ in range(self.alternating_layers // 2)])
        self.local_transformer_blocks.apply(_init_weights)
        self.transformer_input_layernorm = nn.LayerNorm(d, bias=False)
        self.decoder_cross_attn = _CrossBlockP(d, dh)
        self.shared_mlp_output = nn.Sequential(nn.LayerNorm(d), nn.Linear(d, d), nn.GELU(), nn.Linear(d, 3))
        self.shared_mlp_output.apply(_init_weights)

        self._packed = None      # fp16 operand copies of the weights (built lazily on the device)
        self._packed_key = None  # trainable-parameter versions the copies were made from (optimizer steps invalidate them)
        self._packed_frozen = None       # ... the frozen DINOv2 part, rebuilt only on a weight load / device move
        self._packed_frozen_key = None
        self._packed_t = None    # transposed fp16 copies (dgrad operands), training only
        self._train_path = None
        self._ws = {}            # workspace cache
        self._pos_cache = {}
        self.max_decode_rows = 1 << 18

    # ------------------------------------------------------------------ nn.Module plumbing
    def train(self, mode=True):  # Pcd_motion.py:372-373 (returns None, like the reference)
        super().train(mode)

    def load_state_dict(self, *a, **kw):
        self._packed = self._packed_t = self._packed_frozen = None
        if getattr(self, "_graphs", None) is not None:
            self._graphs = {}                     # captured graphs read the packed operand copies, which are re-allocated
        return super().load_state_dict(*a, **kw)

    def _apply(self, fn, *a, **kw):
        self._packed, self._packed_t, self._ws, self._pos_cache, self._train_path = None, None, {}, {}, None
        self._packed_frozen = None
        if getattr(self, "_graphs", None) is not None:
            self._graphs = {}
        return super()._apply(fn, *a, **kw)

    # ------------------------------------------------------------------ weight packing (once per weight load)
    def _buf(self, name, shape, dtype):
        key = (name, tuple(shape), dtype)
        t = self._ws.get(key)
        if t is None:
            t = torch.empty(shape, device=self.pos_embed.device, dtype=dtype)
            self._ws[key] = t
        return t

    def _versions(self):
        return tuple(p._version for p in self.parameters() if p.requires_grad)

    def _frozen_key(self):
        return tuple((p._version, p.data_ptr()) for p in self.image_encoder.parameters())

    def invalidate_packed(self):
        """Drop the fp16 operand copies of the weights (and the CUDA graphs that read them).  They are rebuilt automatically when a
        parameter's version counter moves (optimizer.step(), load_state_dict, in-place ops); writes that bypass the counter
        (``p.data.copy_()``, ``p.data = ...``, EMA swaps) need this call."""
        self._packed = self._packed_t = self._packed_frozen = None
        if getattr(self, "_graphs", None) is not None:
            self._graphs = {}

    def _pack(self):
        """fp16 operand copies of the weights, in two parts: the trainable part is rebuilt when a trainable parameter was modified
        in place since the last call (every optimizer.step()); the frozen DINOv2 part (86 M parameters + the bicubic position table)
        only when one of its tensors was replaced or modified (weight load, .to(), invalidate_packed())."""
        key = self._versions()
        fkey = self._frozen_key()
        if self._packed is not None and self._packed_key == key and self._packed_frozen is not None and self._packed_frozen_key == fkey:
            return self._packed
        dev = self.pos_embed.device

        def w16(weight, kpad=None, split=False):
            w = weight.detach().float().contiguous().reshape(weight.shape[0], -1)
            n, k = w.shape
            kpad = kpad or k
            out = torch.empty(n, kpad * (2 if split else 1), device=dev, dtype=torch.float16)
            ops.cast_pad_f16(w, n, k, out, out.shape[1], kpad, lo_off=kpad if split else 0)
            return out

        def f32(t):
            return t.detach().float().contiguous()

        if self._packed_frozen is None or self._packed_frozen_key != fkey:
            dm = self.image_encoder.model
            Fz = {}
            Fz["d_pe_w"], Fz["d_pe_b"] = w16(dm.patch_embed.proj.weight, KP_PATCH), f32(dm.patch_embed.proj.bias)
            Fz["d_cls"] = f32(dm.cls_token.reshape(-1))
            Fz["d_pos"] = self._dino_pos(dm.pos_embed.detach().float())
            Fz["d_nw"], Fz["d_nb"] = f32(dm.norm.weight), f32(dm.norm.bias)
            Fz["dino"] = [dict(n1w=f32(b.norm1.weight), n1b=f32(b.norm1.bias), qkv=w16(b.attn.qkv.weight), qkv_b=f32(b.attn.qkv.bias),
                               proj=w16(b.attn.proj.weight), proj_b=f32(b.attn.proj.bias), ls1=f32(b.ls1.gamma),
                               n2w=f32(b.norm2.weight), n2b=f32(b.norm2.bias), fc1=w16(b.mlp.fc1.weight), fc1_b=f32(b.mlp.fc1.bias),
                               fc2=w16(b.mlp.fc2.weight), fc2_b=f32(b.mlp.fc2.bias), ls2=f32(b.ls2.gamma)) for b in dm.blocks]
            self._packed_frozen, self._packed_frozen_key = Fz, fkey
        if self._packed is not None and self._packed_key == key:
            self._packed.update(self._packed_frozen)
            return self._packed
        self._packed_t = None
        P = {}
        P["pe_w"], P["pe_b"] = w16(self.point_embed.mlp.weight, KP_EMB, True), f32(self.point_embed.mlp.bias)
        P["pn_w"], P["pn_b"] = w16(self.point_normal_rgb_proj.weight, KP_FEAT, True), f32(self.point_normal_rgb_proj.bias)

        def self_block(m):
            return dict(n1=f32(m.norm1.weight), qkv=w16(m.attn.to_qkv.weight), fc=w16(m.attn.fc.weight),
                        qn=f32(m.attn.q_norm.weight), kn=f32(m.attn.k_norm.weight), n2=f32(m.norm2.weight),
                        w1=w16(m.mlp.mlp[0].weight), w2=w16(m.mlp.mlp[2].weight))

        def cross_block(m):
            kv = torch.empty(2 * self.d, self.d, device=dev, dtype=torch.float16)
            ops.cast_pad_f16(f32(m.attn.to_k.weight), self.d, self.d, kv, self.d, self.d)
            ops.cast_pad_f16(f32(m.attn.to_v.weight), self.d, self.d, kv[self.d:], self.d, self.d)
            return dict(nq=f32(m.norm_q.weight), nkv=f32(m.norm_kv.weight), q=w16(m.attn.to_q.weight), kv=kv,
                        fc=w16(m.attn.fc.weight), qn=f32(m.attn.q_norm.weight), kn=f32(m.attn.k_norm.weight),
                        n2=f32(m.norm2.weight), w1=w16(m.mlp.mlp[0].weight), w2=w16(m.mlp.mlp[2].weight))

        P["enc"], P["dec"] = cross_block(self.encoder_cross_attn), cross_block(self.decoder_cross_attn)
        P["pts"] = [self_block(m) for m in self.points_transformer_blocks]
        P["glb"] = [self_block(m) for m in self.global_transformer_blocks]
        P["loc"] = [self_block(m) for m in self.local_transformer_blocks]
        P["in_ln"] = f32(self.transformer_input_layernorm.weight)
        P["tok"], P["sp0"], P["spr"] = f32(self.learnable_tokens[0]), f32(self.special_token_0[0]), f32(self.special_token_rest[0])
        h = self.shared_mlp_output
        P["h_lnw"], P["h_lnb"] = f32(h[0].weight), f32(h[0].bias)
        P["h1_w"], P["h1_b"] = w16(h[1].weight, self.d, True), f32(h[1].bias)
        P["h3_w"], P["h3_b"] = f32(h[3].weight), f32(h[3].bias)
        P.update(self._packed_frozen)
        self._packed, self._packed_key = P, key
        return P

    def _pack_transposed(self):
        """W^T as fp16 [K_in, N_out] for every trainable nn.Linear: the operand of dX = dY . W (training only)."""
        P = self._pack()
        if self._packed_t is not None:
            return self._packed_t
        dev, d = self.pos_embed.device, self.d

        def wt(weight, rows=None):
            w = weight.detach().float().contiguous()
            n, k = w.shape
            out = torch.empty(k, n, device=dev, dtype=torch.float16)
            ops.cast_transpose_f16(w, n, k, out, n)
            return out

        def self_block(m):
            return dict(qkv=wt(m.attn.to_qkv.weight), fc=wt(m.attn.fc.weight), w1=wt(m.mlp.mlp[0].weight), w2=wt(m.mlp.mlp[2].weight))

        def cross_block(m):
            kv = torch.empty(d, 2 * d, device=dev, dtype=torch.float16)
            ops.cast_transpose_f16(m.attn.to_k.weight.detach().float().contiguous(), d, d, kv, 2 * d)
            ops.cast_transpose_f16(m.attn.to_v.weight.detach().float().contiguous(), d, d, kv[:, d:], 2 * d)
            return dict(q=wt(m.attn.to_q.weight), kv=kv, fc=wt(m.attn.fc.weight), w1=wt(m.mlp.mlp[0].weight), w2=wt(m.mlp.mlp[2].weight))

        PT = dict(enc=cross_block(self.encoder_cross_attn), dec=cross_block(self.decoder_cross_attn),
                  pts=[self_block(m) for m in self.points_transformer_blocks],
                  glb=[self_block(m) for m in self.global_transformer_blocks],
                  loc=[self_block(m) for m in self.local_transformer_blocks],
                  h1=wt(self.shared_mlp_output[1].weight),
                  pn=wt(self.point_normal_rgb_proj.weight))     # [774, 768]; rows 0..767 = the point-embedding columns
        self._packed_t = PT
        return PT

    def _dino_pos(self, pos_embed):
        """Load-time constant folding of DINOv2's position table for the fixed 16x16 grid (upstream
        interpolate_pos_encoding: bicubic, scale factor (16 + 0.1) / 37, class position passed through)."""
        M, n = DINO_GRID, self.hp
        patch = pos_embed[:, 1:].reshape(1, M, M, DINO_DIM).permute(0, 3, 1, 2)
        s = float(n + 0.1) / M
        patch = F.interpolate(patch, scale_factor=(s, s), mode="bicubic", antialias=False)
        assert patch.shape[-2:] == (n, n)
        patch = patch.permute(0, 2, 3, 1).reshape(1, n * n, DINO_DIM)
        return torch.cat([pos_embed[:, :1], patch], dim=1).reshape(-1, DINO_DIM).contiguous()

    def _pos_for(self, T):
        """pos_embed for T input frames; trilinear resize when T != training.frames (Pcd_motion.py:221-228, 481-488).
        A function of a constant buffer and T only, so it is folded once per T."""
        if T == self.latent_length:
            return self.pos_embed.reshape(-1, self.d)
        if T not in self._pos_cache:
            p = self.pos_embed.reshape(1, self.latent_length, self.hp, self.hp, -1).permute(0, 4, 1, 2, 3)
            p = F.interpolate(p, size=(T, self.hp, self.hp), mode="trilinear", align_corners=False)
            self._pos_cache[T] = p.permute(0, 2, 3, 4, 1).reshape(T * self.hp * self.hp, -1).contiguous()
        return self._pos_cache[T]

    # ------------------------------------------------------------------ frame sharding of ONE clip (second row)
    def frame_parallel(self, enabled=True, group=None):
        """Shard the frames of a single clip (B = 1, inference) over the ranks of `group` (default: the world group).  Everything
        per-frame -- DINOv2, token assembly, the 8 local blocks, the decoder, the loss partials -- runs on the rank's own T / W
        frames; only the 8 global blocks couple frames (Pcd_motion.py:401-404): there each rank projects q and k|v for its own
        rows, the k|v rows are all-gathered (one NCCL all-gather of [T*L, 1536] fp16 per global layer, written in place into
        the gathered buffer) and the rank's query rows attend to all keys.  pcd_moved is all-gathered at the end, the loss
        partials are all-reduced.  The shape encoder (64 latent tokens) is replicated."""
        self._fp = (group,) if enabled else None

    def _fp_state(self):
        fp = getattr(self, "_fp", None)
        if fp is None:
            return None
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(fp[0]) == 1:
            return None
        return dist.get_rank(fp[0]), dist.get_world_size(fp[0]), fp[0]

    def _global_block_fp(self, x, rows, w, fp):
        """A global QK_Norm_TransformerBlock over a frame-sharded clip: x holds this rank's rows of the [T*L, d] stream."""
        import torch.distributed as dist
        rank, world, group = fp
        d = self.d
        h = self._buf("h16", (rows, d), torch.float16)
        q = self._buf("fp_q16", (rows, d), torch.float16)
        kv_all = self._buf("fp_kv16", (world * rows, 2 * d), torch.float16)
        o = self._buf("o16", (rows, d), torch.float16)
        hid = self._buf("hid16", (rows, 4 * d), torch.float16)
        kv_loc = kv_all[rank * rows:(rank + 1) * rows]
        ops.layernorm(x, w["n1"], None, 1e-5, rows, d, out16=h, ldo16=d)
        # to_qkv rows [0:768] = Q, [768:2304] = K|V (transformer.py:200): two GEMMs so that K|V lands contiguously in its
        # slice of the gathered buffer (in-place all-gather, no pack / unpack copies)
        ops.gemm(h, w["qkv"][d:], rows, 2 * d, d, out16=kv_loc, ldo16=2 * d, qn_w=w["kn"], kn_w=None, qk_eps=1e-5, qk_cols=d)
        # fp_overlap (opt-in): hide the gather behind the attention over the rank's own keys.  Measured on 2 and 4 B200s
        # (profiles/r1p_frame_shard_overlap.jsonl) the NVLink all-gather is too cheap for that to pay: the three shorter
        # launches + the merge cost more than the gather they hide (128 frames, 4 GPUs: 25.8 ms vs 23.6 ms blocking).
        overlap = getattr(self, "fp_overlap", False)
        if overlap:   # issued right behind the K|V projection: the gather travels while q is projected and the own keys are attended
            work = dist.all_gather_into_tensor(kv_all, kv_loc, group=group, async_op=True)
        ops.gemm(h, w["qkv"], rows, d, d, out16=q, ldo16=d, qn_w=w["qn"], kn_w=None, qk_eps=1e-5, qk_cols=d)
        kw = dict(B=1, H=self.H, Lq=rows, q_ld=d, k_ld=2 * d, v_ld=2 * d, o_ld=d, q_rows=rows, q_batch_rows=rows, scale=self.dh ** -0.5)
        if not overlap:
            dist.all_gather_into_tensor(kv_all, kv_loc, group=group)
            ops.attention(q, kv_all, kv_all[:, d:], o, Lk=world * rows, kv_rows=world * rows, kv_batch_rows=world * rows, **kw)
        else:
            # Overlap: the all-gather runs on NCCL's stream while this rank attends to the keys it already owns; the gathered
            # keys (the rows before and after its own slice) follow as further partial launches over the same queries, and
            # m324_attention_merge combines the (O, m, l) triples (log-sum-exp).
            ranges = [(rank * rows, rows)]
            if rank > 0:
                ranges.append((0, rank * rows))
            if rank < world - 1:
                ranges.append(((rank + 1) * rows, (world - 1 - rank) * rows))
            parts = len(ranges)
            ws = self._buf("fp_attn_ws", (ops.attention_partial_bytes(1, self.H, rows, 3),), torch.uint8)
            for idx, (r0, ln) in enumerate(ranges):
                if idx == 1:
                    work.wait()       # the compute stream waits for the gather only now
                ops.attention(q, kv_all[r0:], kv_all[r0:, d:], o, Lk=ln, kv_rows=ln, kv_batch_rows=ln, workspace=ws, partial=(parts, idx), **kw)
            ops.attention_merge(o, B=1, H=self.H, Lq=rows, o_ld=d, parts=parts, workspace=ws)
        ops.gemm(o, w["fc"], rows, d, d, resid=x, ldr=d, out32=x, ldo32=d)
        ops.layernorm(x, w["n2"], None, 1e-5, rows, d, out16=h, ldo16=d)
        ops.gemm(h, w["w1"], rows, 4 * d, d, act=1, out16=hid, ldo16=4 * d)
        ops.gemm(hid, w["w2"], rows, d, 4 * d, resid=x, ldr=d, out32=x, ldo32=d)

    # ------------------------------------------------------------------ kernel-launch helpers
    def _self_block(self, x, rows, Batt, L, w, tag):
        """QK_Norm_TransformerBlock.forward (transformer.py:420-423) on the fp32 residual stream x [rows, d], in place."""
        d = self.d
        h = self._buf("h16", (rows, d), torch.float16)
        qkv = self._buf("qkv16", (rows, 3 * d), torch.float16)
        o = self._buf("o16", (rows, d), torch.float16)
        hid = self._buf("hid16", (rows, 4 * d), torch.float16)
        ops.layernorm(x, w["n1"], None, 1e-5, rows, d, out16=h, ldo16=d)
        ops.gemm(h, w["qkv"], rows, 3 * d, d, out16=qkv, ldo16=3 * d, qn_w=w["qn"], kn_w=w["kn"], qk_eps=1e-5, qk_cols=d)
        ops.attention(qkv, qkv[:, d:], qkv[:, 2 * d:], o, B=Batt, H=self.H, Lq=L, Lk=L, q_ld=3 * d, k_ld=3 * d, v_ld=3 * d,
                      o_ld=d, q_rows=rows, kv_rows=rows, q_batch_rows=L, kv_batch_rows=L, scale=self.dh ** -0.5)
        ops.gemm(o, w["fc"], rows, d, d, resid=x, ldr=d, out32=x, ldo32=d)
        ops.layernorm(x, w["n2"], None, 1e-5, rows, d, out16=h, ldo16=d)
        ops.gemm(h, w["w1"], rows, 4 * d, d, act=1, out16=hid, ldo16=4 * d)
        ops.gemm(hid, w["w2"], rows, d, 4 * d, resid=x, ldr=d, out32=x, ldo32=d)

    def _point_features(self, P, xyz, normal, rgb, n, out32):
        """point_normal_rgb_proj(cat[point_embed(xyz), normal, rgb]) (Pcd_motion.py:456-459, 550-553) -> fp32 [n, d]."""
        d = self.d
        a0 = self._buf("pf_a0", (n, 2 * KP_EMB), torch.float16)
        a1 = self._buf("pf_a1", (n, 2 * KP_FEAT), torch.float16)
        ops.point_embed_features(xyz, n, a0, 2 * KP_EMB, KP_EMB)
        ops.gemm(a0, P["pe_w"], n, d, KP_EMB, passes=3, a_lo_off=KP_EMB, w_lo_off=KP_EMB, bias=P["pe_b"], out16=a1,
                 ldo16=2 * KP_FEAT, out16_lo_off=KP_FEAT)
        ops.point_extra_features(normal, rgb, n, a1, 2 * KP_FEAT, d, KP_FEAT, KP_FEAT)
        ops.gemm(a1, P["pn_w"], n, d, KP_FEAT, passes=3, a_lo_off=KP_FEAT, w_lo_off=KP_FEAT, bias=P["pn_b"], out32=out32, ldo32=d)

    def _side_stream(self):
        dev = self.pos_embed.device
        if getattr(self, "_side", None) is None or self._side.device != dev:
            self._side = torch.cuda.Stream(device=dev)
        return self._side

    def _shape_encoder(self, P, sample, B, S, M):
        """Pcd_motion.py:456-464: point features -> cross-attention into the learnable tokens -> 4 self-attention blocks."""
        d, H, dh = self.d, self.H, self.dh
        scale = dh ** -0.5
        f32c = lambda t: t.detach().float().contiguous()
        shape_feat = self._buf("shape_feat", (B * S, d), torch.float32)
        self._point_features(P, f32c(sample["ref_shape_pcd"]).reshape(-1, 3), f32c(sample["ref_shape_normals"]).reshape(-1, 3),
                             f32c(sample["ref_shape_rgbs"]).reshape(-1, 3), B * S, shape_feat)
        mesh = self._buf("mesh_feat", (B * M, d), torch.float32)
        e = P["enc"]
        qn16 = self._buf("enc_qn16", (M, d), torch.float16)
        q16 = self._buf("enc_q16", (M, d), torch.float16)
        kn16 = self._buf("enc_kn16", (B * S, d), torch.float16)
        kv16 = self._buf("enc_kv16", (B * S, 2 * d), torch.float16)
        eo16 = self._buf("enc_o16", (B * M, d), torch.float16)
        ops.layernorm(P["tok"], e["nq"], None, 1e-5, M, d, out16=qn16, ldo16=d)
        ops.gemm(qn16, e["q"], M, d, d, out16=q16, ldo16=d, qn_w=e["qn"], kn_w=None, qk_cols=d)
        ops.layernorm(shape_feat, e["nkv"], None, 1e-5, B * S, d, out16=kn16, ldo16=d)
        ops.gemm(kn16, e["kv"], B * S, 2 * d, d, out16=kv16, ldo16=2 * d, qn_w=e["kn"], kn_w=None, qk_cols=d)
        ops.attention(q16, kv16, kv16[:, d:], eo16, B=B, H=H, Lq=M, Lk=S, q_ld=d, k_ld=2 * d, v_ld=2 * d, o_ld=d, q_rows=M,
                      kv_rows=B * S, q_batch_rows=0, kv_batch_rows=S, scale=scale)
        ops.gemm(eo16, e["fc"], B * M, d, d, resid=P["tok"], ldr=d, resid_mod=M, out32=mesh, ldo32=d)
        h16 = self._buf("h16s", (B * M, d), torch.float16)
        hid16 = self._buf("hid16s", (B * M, 4 * d), torch.float16)
        ops.layernorm(mesh, e["n2"], None, 1e-5, B * M, d, out16=h16, ldo16=d)
        ops.gemm(h16, e["w1"], B * M, 4 * d, d, act=1, out16=hid16, ldo16=4 * d)
        ops.gemm(hid16, e["w2"], B * M, d, 4 * d, resid=mesh, ldr=d, out32=mesh, ldo32=d)
        for w in P["pts"]:
            self._self_block(mesh, B * M, B, M, w, "pts")
        return mesh


    def _dino_forward(self, P, rgb_video, Fr, Hin, Win):
        """Frozen DINOv2 ViT-B/14 per frame (Pcd_motion.py:466-475, image_encoder/dinov2.py:65-124) -> fp32 residual stream
        [Fr * 257, d] before the final norm (the norm is fused into the token assembly)."""
        d, H = self.d, self.H
        scale = self.dh ** -0.5
        npatch = self.hp * self.hp
        patches = self._buf("patches", (Fr * npatch, KP_PATCH), torch.float16)
        ops.preprocess_frames(rgb_video.detach().float().contiguous(), Fr, Hin, Win, self.image_size, patches, KP_PATCH, KP_PATCH)
        pe_out = self._buf("patch_embed", (Fr * npatch, d), torch.float32)
        ops.gemm(patches, P["d_pe_w"], Fr * npatch, d, KP_PATCH, bias=P["d_pe_b"], out32=pe_out, ldo32=d)
        Ld = npatch + 1
        rows_d = Fr * Ld
        xd = self._buf("dino_x", (rows_d, d), torch.float32)
        ops.dino_assemble(pe_out, P["d_cls"], P["d_pos"], Fr, npatch, d, xd)
        hd = self._buf("h16", (rows_d, d), torch.float16)
        qkvd = self._buf("qkv16", (rows_d, 3 * d), torch.float16)
        od = self._buf("o16", (rows_d, d), torch.float16)
        hidd = self._buf("hid16", (rows_d, 4 * d), torch.float16)
        for w in P["dino"]:
            ops.layernorm(xd, w["n1w"], w["n1b"], DINO_EPS, rows_d, d, out16=hd, ldo16=d)
            ops.gemm(hd, w["qkv"], rows_d, 3 * d, d, bias=w["qkv_b"], out16=qkvd, ldo16=3 * d)
            ops.attention(qkvd, qkvd[:, d:], qkvd[:, 2 * d:], od, B=Fr, H=H, Lq=Ld, Lk=Ld, q_ld=3 * d, k_ld=3 * d, v_ld=3 * d,
                          o_ld=d, q_rows=rows_d, kv_rows=rows_d, q_batch_rows=Ld, kv_batch_rows=Ld, scale=scale)
            ops.gemm(od, w["proj"], rows_d, d, d, bias=w["proj_b"], gamma=w["ls1"], resid=xd, ldr=d, out32=xd, ldo32=d)
            ops.layernorm(xd, w["n2w"], w["n2b"], DINO_EPS, rows_d, d, out16=hd, ldo16=d)
            ops.gemm(hd, w["fc1"], rows_d, 4 * d, d, bias=w["fc1_b"], act=1, out16=hidd, ldo16=4 * d)
            ops.gemm(hidd, w["fc2"], rows_d, d, 4 * d, bias=w["fc2_b"], gamma=w["ls2"], resid=xd, ldr=d, out32=xd, ldo32=d)
        return xd, npatch

    # ------------------------------------------------------------------ training
    def trainable_parameters(self):
        return [(n, p) for n, p in self.named_parameters() if p.requires_grad]

    def grad_buffer(self):
        """The flat fp32 gradient buffer (model/train_path.py:GradBuffer): ``.flat`` is what a data-parallel trainer
        all-reduces (one ncclAllReduce, train.py's DDP C1), ``.views[name]`` the per-parameter slices."""
        if self._train_path is None:
            self._train_path = TrainPath(self)
        return self._train_path.grad_buffer()

    @torch.no_grad()
    def forward_backward(self, sample, zero_grads=True, grad_scale=1.0, allreduce_group=False):
        """One training forward + backward on libm324 (what train.py:157-170 does with autocast + loss.backward()).
        Gradients of ``grad_scale * loss`` are accumulated into grad_buffer().flat and every trainable parameter's
        ``.grad`` is set to its slice (no copy), ready for ``all_reduce(flat)`` and ``optimizer.step()``.
        ``allreduce_group`` (None = the world group, a ProcessGroup, or False = off): average the gradients over the data-parallel
        ranks INSIDE this call, overlapped with the backward (train_path.OverlappedAllReduce: the exchange step DDP gives
        train.py:88-89, in three waves behind the compute stream); the returned loss metrics are then the averaged ones."""
        if not sample["ref_pcd"].is_cuda:
            raise RuntimeError("Motion_Latent_Model (libm324) runs on CUDA tensors only: there is no CPU path")
        if sample["ref_pcd"].device.index != torch.cuda.current_device():
            with torch.cuda.device(sample["ref_pcd"].device):
                return self.forward_backward(sample, zero_grads, grad_scale, allreduce_group)
        gb = self.grad_buffer()
        ar = None
        if allreduce_group is not False:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size(allreduce_group) > 1:
                from .train_path import OverlappedAllReduce
                ar = OverlappedAllReduce(gb, allreduce_group)
        self._train_path.on_ready = ar.ready if ar is not None else None
        try:
            out, loss = self._train_path.run(sample, zero_grads=zero_grads, grad_scale=grad_scale)
        finally:
            self._train_path.on_ready = None
        if ar is not None:
            loss = ar.finish()
        loss = loss.clone()      # the live copy sits in the tail of the flat gradient buffer (averaged by allreduce_gradients)
        for n, p in self.trainable_parameters():
            p.grad = gb.views[n]
        lm = edict()
        lm.loss, lm.xyz_loss = loss[1], loss[0]
        return edict(input_data=sample, pcd_moved=out, loss_metrics=lm)

    def allreduce_gradients(self, group=None):
        """Average the gradients (and the loss metrics of the last forward_backward) over the data-parallel ranks with ONE
        all-reduce of the flat gradient buffer -- the exchange step train.py gets from DDP (train.py:88-89); every
        parameter's ``.grad`` is a view of that buffer, so ``optimizer.step()`` follows directly.  Returns
        edict(loss, xyz_loss) averaged over ranks."""
        m = self.grad_buffer().allreduce(group)
        lm = edict()
        lm.loss, lm.xyz_loss = m[1], m[0]
        return lm

    def _forward_autograd(self, sample):
        """train() + grad enabled: the loss gets a grad_fn so that train.py:162 (``loss.backward()``), DDP's reducer hooks
        and gradient accumulation behave as with the reference."""
        names = [n for n, _ in self.trainable_parameters()]
        params = [p for _, p in self.trainable_parameters()]
        out, loss, xyz = _TrainStepFn.apply(self, sample, names, *params)
        lm = edict()
        lm.loss, lm.xyz_loss = loss, xyz
        return edict(input_data=sample, pcd_moved=out, loss_metrics=lm)

    def _warn_once(self, msg):
        if not getattr(self, "_warned", False):
            import warnings
            warnings.warn(msg, stacklevel=3)
            self._warned = True

    # ------------------------------------------------------------------ CUDA-graph replay of the inference forward
    def enable_cuda_graph(self, enabled=True):
        """Replay the inference forward from a CUDA graph (one graph per input-shape signature, captured on first use).  Every
        launch of the forward is capture-safe by construction (include/m324.h: no allocation, no synchronisation, TMA
        descriptors passed by value as kernel parameters; workspaces are cached per shape, so their addresses are static);
        inputs are copied into static device buffers before each replay.  The returned tensors are the graph's static outputs:
        they are overwritten by the next call with the same shapes (clone them to keep them).  Not used in train() mode."""
        self._graphs = {} if enabled else None

    def _graph_forward(self, sample):
        tens = {k: v for k, v in sample.items() if torch.is_tensor(v)}
        key = tuple(sorted((k, tuple(v.shape), v.dtype) for k, v in tens.items()))
        if self._packed is None or self._packed_key != self._versions() or self._packed_frozen is None or self._packed_frozen_key != self._frozen_key():
            self._graphs.clear()        # weights changed in place (optimizer step): the packed copies the graphs read are stale
        entry = self._graphs.get(key)
        if entry is None:
            static = {k: v.detach().clone() for k, v in tens.items()}
            self._pack()
            cap = torch.cuda.Stream(device=self.pos_embed.device)
            cap.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(cap):
                for _ in range(2):      # warm-up on the capture stream: workspaces, the attention scratch and position tables get allocated here
                    self._forward_inference(static)
            torch.cuda.current_stream().wait_stream(cap)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=cap):
                out = self._forward_inference(static)
            entry = self._graphs[key] = (graph, static, out)
        graph, static, out = entry
        for k, v in tens.items():
            static[k].copy_(v, non_blocking=True)
        graph.replay()
        res = edict(input_data=sample, pcd_moved=out.pcd_moved)
        if "loss_metrics" in out:
            res.loss_metrics = out.loss_metrics
        return res

    # ------------------------------------------------------------------ forward
    def forward(self, sample):
        ref = sample["ref_pcd"]
        if ref.is_cuda and ref.device.index != torch.cuda.current_device():
            # the library launches on the CURRENT device's stream (ops._stream) and keeps per-device kernel attributes:
            # run on the device the tensors live on, like the reference module does after .to(device)
            with torch.cuda.device(ref.device):
                return self._forward(sample)
        return self._forward(sample)

    def _forward(self, sample):
        if self.training and torch.is_grad_enabled() and "point_clouds" not in sample:
            self._warn_once("train() mode with grad enabled but no 'point_clouds' in the sample: the reference returns a differentiable "
                            "pcd_moved here; this build runs its inference forward (no grad_fn) -- pass the targets to train")
        if self.training and torch.is_grad_enabled() and "point_clouds" in sample:
            if not sample["ref_pcd"].is_cuda:
                raise RuntimeError("Motion_Latent_Model (libm324) runs on CUDA tensors only: there is no CPU path")
            if self._fp_state() is not None:
                raise RuntimeError("frame_parallel() shards the frames of one clip for inference; training shards clips (train.py:58-59)")
            return self._forward_autograd(sample)
        with torch.no_grad():
            if getattr(self, "_graphs", None) is not None and not self.training and self._fp_state() is None and sample["ref_pcd"].is_cuda:
                return self._graph_forward(sample)
            return self._forward_inference(sample)

    def _forward_inference(self, sample):
        ref_pcd = sample["ref_pcd"]
        if not ref_pcd.is_cuda:
            raise RuntimeError("Motion_Latent_Model (libm324) runs on CUDA tensors only: there is no CPU path")
        P = self._pack()
        d, H, dh = self.d, self.H, self.dh
        scale = dh ** -0.5
        f32c = lambda t: t.detach().float().contiguous()
        B, N = ref_pcd.shape[:2]
        S = sample["ref_shape_pcd"].shape[1]
        M = self.num_learnable_tokens
        rgb_video = sample["rgb_video"]
        T_all, Hin, Win = rgb_video.shape[1:4]
        fp = self._fp_state()
        t_first = 0
        T = T_all
        if fp is not None:      # this rank's frames of the single clip
            if B != 1:
                raise ValueError(f"frame_parallel shards the frames of ONE clip (B = 1); got B={B}")
            t_first, T = frame_shard(T_all, fp[0], fp[1])
            rgb_video = rgb_video[:, t_first:t_first + T]
        Fr = B * T

        # ---- A. shape encoder (Pcd_motion.py:456-464).  Independent of the video branch until token assembly: it is a chain
        # of ~40 tiny launches (64 latent tokens), so it runs on a side stream underneath the DINOv2 kernels.
        main_stream = torch.cuda.current_stream()
        side = self._side_stream()
        side.wait_stream(main_stream)
        with torch.cuda.stream(side):
            mesh = self._shape_encoder(P, sample, B, S, M)
        shape_done = torch.cuda.Event()
        shape_done.record(side)

        # ---- B. frozen DINOv2 ViT-B/14 per frame (Pcd_motion.py:466-475, image_encoder/dinov2.py:65-124)
        xd, npatch = self._dino_forward(P, rgb_video, Fr, Hin, Win)

        # ---- C. token assembly + transformer_input_layernorm (Pcd_motion.py:477-509)
        L = 4 + M + npatch
        rows_t = Fr * L
        x = self._buf("trunk_x", (rows_t, d), torch.float32)
        main_stream.wait_event(shape_done)
        # pos_drop (Pcd_motion.py:369-370, 490) is active whenever the module is in train() mode, also under no_grad
        drop_p = float(self.drop_rate) if self.training else 0.0
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if drop_p > 0 else 0
        pos = self._pos_for(T_all)
        if fp is not None:      # rows of this rank's frames; only the clip's very first frame carries special_token_0 (:495-500)
            pos = pos[t_first * npatch:(t_first + T) * npatch]
        ops.assemble_tokens(xd, P["d_nw"], P["d_nb"], DINO_EPS, pos, P["sp0"] if t_first == 0 else P["spr"], P["spr"], mesh,
                            P["in_ln"], 1e-5, B, T, M, npatch, d, x, drop_p=drop_p, seed=seed + t_first)

        # ---- D. alternating global / local attention (Pcd_motion.py:394-409)
        for wg, wl in zip(P["glb"], P["loc"]):
            if fp is not None:
                self._global_block_fp(x, rows_t, wg, fp)
            else:
                self._self_block(x, rows_t, B, T * L, wg, "glb")
            self._self_block(x, rows_t, Fr, L, wl, "loc")

        # ---- E. per-frame cross-attention decoder + output head + loss (Pcd_motion.py:520-579, model/loss.py:59-61)
        dc = P["dec"]
        feat = self._buf("dec_feat", (B * N, d), torch.float32)
        self._point_features(P, f32c(ref_pcd).reshape(-1, 3), f32c(sample["ref_normal"]).reshape(-1, 3),
                             f32c(sample["ref_rgb"]).reshape(-1, 3), B * N, feat)
        dqn16 = self._buf("dec_qn16", (B * N, d), torch.float16)
        dq16 = self._buf("dec_q16", (B * N, d), torch.float16)
        ops.layernorm(feat, dc["nq"], None, 1e-5, B * N, d, out16=dqn16, ldo16=d)
        ops.gemm(dqn16, dc["q"], B * N, d, d, out16=dq16, ldo16=d, qn_w=dc["qn"], kn_w=None, qk_cols=d)
        dkn16 = self._buf("dec_kn16", (Fr * M, d), torch.float16)
        dkv16 = self._buf("dec_kv16", (Fr * M, 2 * d), torch.float16)
        # exact token slice [:, :, 4:4+tokens] (Pcd_motion.py:520) as a row gather inside the LayerNorm
        ops.layernorm(x, dc["nkv"], None, 1e-5, Fr * M, d, src_rpg=M, src_gstride=L, src_goff=4, out16=dkn16, ldo16=d)
        ops.gemm(dkn16, dc["kv"], Fr * M, 2 * d, d, out16=dkv16, ldo16=2 * d, qn_w=dc["kn"], kn_w=None, qk_cols=d)

        out_all = torch.empty(B, T_all, N, 3, device=ref_pcd.device, dtype=torch.float32)
        out = out_all[:, t_first:t_first + T]       # this rank's frames (the whole tensor without frame sharding)
        target = f32c(sample["point_clouds"]) if "point_clouds" in sample else None
        if target is not None and tuple(target.shape) != (B, T_all, N, 3):  # model/loss.py:50-57
            raise ValueError("Shape mismatch or invalid shape for coordinate MSE. Expected both tensors of shape (B, T, N, C). "
                             f"Got pred: {tuple(out_all.shape)}, target: {tuple(target.shape)}")
        if target is not None:
            target = target[:, t_first:t_first + T]
        weight = float(self.config.training.coord_mse_loss_weight)
        partials = self._buf("mse_partials", (Fr * 1024,), torch.float32) if target is not None else None
        n_part = 0
        tchunk = max(1, min(T, self.max_decode_rows // max(N, 1)))
        for b in range(B):
            for t0 in range(0, T, tchunk):
                tc = min(tchunk, T - t0)
                rows = tc * N
                f0 = b * T + t0
                o16 = self._buf("dec_o16", (tchunk * N, d), torch.float16)
                xdec = self._buf("dec_x", (tchunk * N, d), torch.float32)
                dh16 = self._buf("dec_h16", (tchunk * N, 2 * d), torch.float16)
                dhid = self._buf("dec_hid16", (tchunk * N, 4 * d), torch.float16)
                hpart = self._buf("dec_hpart", (tchunk * N, d // 64, 4), torch.float32)
                ops.attention(dq16[b * N:], dkv16[f0 * M:], dkv16[f0 * M:, d:], o16, B=tc, H=H, Lq=N, Lk=M, q_ld=d, k_ld=2 * d,
                              v_ld=2 * d, o_ld=d, q_rows=N, kv_rows=tc * M, q_batch_rows=0, kv_batch_rows=M, scale=scale)
                ops.gemm(o16, dc["fc"], rows, d, d, resid=feat[b * N:], ldr=d, resid_mod=N, out32=xdec, ldo32=d)
                ops.layernorm(xdec, dc["n2"], None, 1e-5, rows, d, out16=dh16, ldo16=2 * d)
                ops.gemm(dh16, dc["w1"], rows, 4 * d, d, lda=2 * d, act=1, out16=dhid, ldo16=4 * d)
                ops.gemm(dhid, dc["w2"], rows, d, 4 * d, resid=xdec, ldr=d, out32=xdec, ldo32=d)
                # shared_mlp_output (Pcd_motion.py:336-341, 561): LN(+bias) -> Linear+GELU (split fp16) -> Linear(768,3) in fp32
                ops.layernorm(xdec, P["h_lnw"], P["h_lnb"], 1e-5, rows, d, out16=dh16, ldo16=2 * d, lo_off=d)
                # ... with the 768 -> 3 projection folded into that GEMM's epilogue: the [rows, 768] fp32 hidden tensor is never written
                ops.gemm(dh16, P["h1_w"], rows, d, d, passes=3, a_lo_off=d, w_lo_off=d, bias=P["h1_b"], act=1, head_w=P["h3_w"], head_part=hpart)
                o_view = out[b, t0:t0 + tc]
                tgt = target[b, t0:t0 + tc] if target is not None else None
                n_part += ops.head3_from_partials(hpart, d // 64, P["h3_b"], rows, o_view, tgt,
                                                  partials[n_part:] if partials is not None else None)

        if fp is not None:
            import torch.distributed as dist
            dist.all_gather_into_tensor(out_all.view(-1), out.reshape(-1), group=fp[2])     # in place: rank r owns frames [r*T, (r+1)*T)
        result = edict(input_data=sample, pcd_moved=out_all)
        if target is not None:
            loss = torch.empty(2, device=ref_pcd.device, dtype=torch.float32)
            ops.mse_finalize(partials, n_part, float(B) * T_all * N * 3, weight, loss)    # this rank's share of the clip's mean
            if fp is not None:
                dist.all_reduce(loss, group=fp[2])
            lm = edict()
            lm.loss = loss[1]
            lm.xyz_loss = loss[0]
            result.loss_metrics = lm
        return result


class _TrainStepFn(torch.autograd.Function):
    """Autograd seam of the training step: forward() runs libm324's forward AND backward (the decoder is differentiated
    chunk by chunk while its activations are live), backward() hands the stored parameter gradients to autograd."""

    @staticmethod
    def forward(ctx, model, sample, names, *params):
        if model._train_path is None:
            model._train_path = TrainPath(model)
        tp = model._train_path
        out, loss = tp.run(sample, zero_grads=True, grad_scale=1.0)
        loss = loss.clone()
        ctx.model, ctx.names, ctx.step_id = model, names, tp.step_id
        ctx.mark_non_differentiable(out)
        return out, loss[1], loss[0]

    @staticmethod
    def backward(ctx, g_out, g_loss, g_xyz):
        tp = ctx.model._train_path
        if tp.step_id != ctx.step_id:
            raise RuntimeError("Motion_Latent_Model: backward() of a stale forward -- the gradient buffer was overwritten by a later "
                               "training forward; call loss.backward() before the next model(batch)")
        if getattr(tp, "scaled_step", -1) == ctx.step_id:
            raise RuntimeError("Motion_Latent_Model: a second backward() through the same training forward (retain_graph) is not "
                               "supported -- the gradients live in one flat buffer that the first backward() has already scaled")
        tp.scaled_step = ctx.step_id
        gb = tp.grad_buffer()
        # d(total) / d(params) = (g_loss + g_xyz / weight) * d(loss) / d(params)   (xyz_loss = loss / weight, model/loss.py:59-61);
        # the scalars are read on the device: no host synchronisation (train.py:159-166 passes loss / grad_accum_steps)
        w = float(ctx.model.config.training.coord_mse_loss_weight)
        ga = g_loss.detach().float().reshape(()).contiguous() if g_loss is not None else None
        gx = g_xyz.detach().float().reshape(()).contiguous() if (g_xyz is not None and w != 0.0) else None
        if ga is None and gx is None:
            gb.flat[:gb.n_grad].zero_()
        else:
            ops.scale_by_device_scalars(gb.flat, gb.n_grad, ga, gx, (1.0 / w) if gx is not None else 0.0)
        return (None, None, None) + tuple(gb.views[n] for n in ctx.names)
