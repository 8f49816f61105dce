"""ORACLE SUPPORT (test infrastructure): make the UNMODIFIED reference importable in the build
container, where three of its imports are missing:

 1. ``easydict``            -> a dict subclass with attribute access;
 2. ``xformers.ops``        -> ``memory_efficient_attention`` = exact softmax attention in the
                               input dtype, layout [B, L, H, Dh] (transformer.py:134-139, 209-214);
 3. ``torch.hub.load``      -> returns oracle.dinov2_vitb14.DinoV2ViTB14 (dinov2.py:44 needs the
                               network otherwise).

The reference tree is /root/reference in the build container and its byte-identical staged copy oracle/_ref on the
GPU box (oracle/build_ref.py, SHA-256 manifest).  Users: tests/golden/make_golden*.py (fixture generation), bench.py's
``--impl reference`` arm (the unmodified reference on the host cores) and its ``reference_gpu`` leg (the unmodified
reference on the B200 under bf16 autocast with ``attention="flash"``: xformers' flash op -> flash_attn.flash_attn_func,
SURVEY.md 8(c)/(d)), scripts/run_reference_train.py.  The product path never imports this.
"""
import sys
import types

import torch

from . import build_ref


def reference_root():
    return build_ref.root()


class EasyDict(dict):
    def __init__(self, d=None, **kw):
        super().__init__()
        d = dict(d or {}, **kw)
        for k, v in d.items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, EasyDict):
            v = EasyDict(v)
        super().__setitem__(k, v)

    __setattr__ = __setitem__

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


def _mea(q, k, v, attn_bias=None, p=0.0, op=None, scale=None):
    assert attn_bias is None and p == 0.0
    scale = q.shape[-1] ** -0.5 if scale is None else scale
    q_, k_, v_ = (t.transpose(1, 2) for t in (q, k, v))
    B, H, Lq, _ = q_.shape
    Lk = k_.shape[2]
    rows = max(1, min(Lq, (1 << 28) // max(1, B * H * Lk)))     # score blocks of <= 2^28 elements (the chili clip has 51,516 tokens)
    if rows >= Lq:
        s = (q_ @ k_.transpose(-2, -1)) * scale
        return (torch.softmax(s, dim=-1) @ v_).transpose(1, 2)
    out = torch.empty_like(q_)
    kt = k_.transpose(-2, -1)
    for r0 in range(0, Lq, rows):                                # exact: softmax rows are independent
        s = (q_[:, :, r0:r0 + rows] @ kt) * scale
        out[:, :, r0:r0 + rows] = torch.softmax(s, dim=-1) @ v_
    return out.transpose(1, 2)


def _mea_flash(q, k, v, attn_bias=None, p=0.0, op=None, scale=None):
    """What xformers' fmha.flash.FwOp dispatches to (transformer.py:134-139, 209-214): flash-attn 2, layout [B, L, H, Dh],
    fp16 / bf16 only.  fp32 inputs (no autocast) are rejected by xformers' flash op; here they are rounded to bf16 so that a
    caller outside autocast still runs."""
    from flash_attn import flash_attn_func
    assert attn_bias is None and p == 0.0
    dt = q.dtype
    if dt not in (torch.float16, torch.bfloat16):
        q, k, v = (t.to(torch.bfloat16) for t in (q, k, v))
    return flash_attn_func(q, k, v, dropout_p=0.0, softmax_scale=scale, causal=False).to(dt)


def install(attention="exact"):
    """attention: "exact" (fp32-capable softmax attention, CPU or GPU) or "flash" (flash_attn_func, GPU bf16/fp16)."""
    if "easydict" not in sys.modules:
        m = types.ModuleType("easydict")
        m.EasyDict = EasyDict
        sys.modules["easydict"] = m
    if "xformers" in sys.modules and getattr(sys.modules["xformers"], "_m324_shim", False):
        sys.modules["xformers.ops"].memory_efficient_attention = _mea_flash if attention == "flash" else _mea
    if "xformers" not in sys.modules:
        xf = types.ModuleType("xformers")
        xf._m324_shim = True
        ops = types.ModuleType("xformers.ops")
        fmha = types.ModuleType("xformers.ops.fmha")
        flash = types.ModuleType("xformers.ops.fmha.flash")
        flash.FwOp = object()
        flash.BwOp = object()
        fmha.flash = flash
        ops.fmha = fmha
        ops.memory_efficient_attention = _mea_flash if attention == "flash" else _mea
        xf.ops = ops
        sys.modules.update({"xformers": xf, "xformers.ops": ops, "xformers.ops.fmha": fmha,
                            "xformers.ops.fmha.flash": flash})
    from . import dinov2_vitb14

    def _hub_load(repo, name, *a, **kw):
        assert repo == "facebookresearch/dinov2" and name == "dinov2_vitb14", (repo, name)
        return dinov2_vitb14.DinoV2ViTB14()

    torch.hub.load = _hub_load
    dinov2_vitb14.FUSED_SDPA = attention == "flash"   # upstream MemEffAttention = a fused kernel on the GPU
    root = reference_root()
    if root not in sys.path:
        sys.path.insert(0, root)


def install_trimesh_stub():
    """``from dataset.dyscene import collate_fn_with_topology`` (train.py:18) imports trimesh at module level and names
    ``trimesh.Trimesh`` in an annotation; the collate function itself never touches it.  An attribute-only stand-in."""
    if "trimesh" not in sys.modules:
        m = types.ModuleType("trimesh")

        class Trimesh:      # noqa: D401  (annotation target only)
            def __init__(self, *a, **kw):
                raise RuntimeError("trimesh is not installed in this image (stand-in for the import only)")

        m.Trimesh = Trimesh
        sys.modules["trimesh"] = m


def make_config(frames=12, drop_rate=0.0, use_checkpoint=False):
    """configs/dyscene.yaml as an EasyDict (model + the training keys forward() reads)."""
    return EasyDict({
        "model": {"class_name": "model.Pcd_motion.Motion_Latent_Model", "feat_dim": 768, "tokens": 64,
                  "pcd_layers": 4,
                  "video_encoder": {"image_tokenizer": {"image_size": 224, "patch_size": 14, "patch_length": 1,
                                                        "in_channels": 3},
                                    "transformer": {"d": 768, "d_head": 64, "n_layer": 16, "special_init": True,
                                                    "depth_init": True, "use_qk_norm": True,
                                                    "drop_rate": drop_rate}}},
        "training": {"frames": frames, "use_checkpoint": use_checkpoint, "grad_checkpoint_every": 1,
                     "coord_mse_loss_weight": 1.0, "amp_dtype": "bf16", "use_amp": True, "use_tf32": True},
    })


def build_reference_model(frames=12, attention="exact"):
    """Construct the unmodified reference Motion_Latent_Model (eval mode)."""
    install(attention)
    import importlib
    mod = importlib.import_module("model.Pcd_motion")
    model = mod.Motion_Latent_Model(make_config(frames=frames))
    model.eval()
    return model
