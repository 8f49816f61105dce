"""Cross-check the oracle's DINOv2 ViT-B/14 restatement (un-vendored third party of the reference, parity otherwise
unpinned) against the independent `transformers` Dinov2 implementation with the same random weights.  CPU only."""
import pytest
import torch

from oracle import dinov2_vitb14 as dino
from oracle import motion324_oracle as orc

transformers = pytest.importorskip("transformers")


def _hf_model_from(sd, pfx="image_encoder.model."):
    from transformers import Dinov2Config, Dinov2Model
    cfg = Dinov2Config(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, mlp_ratio=4, image_size=518, patch_size=14,
                       layer_norm_eps=1e-6, layerscale_value=1.0, qkv_bias=True, hidden_act="gelu", use_swiglu_ffn=False)
    m = Dinov2Model(cfg).eval()
    hs = m.state_dict()
    g = lambda k: sd[pfx + k]
    hs["embeddings.cls_token"] = g("cls_token")
    hs["embeddings.mask_token"] = g("mask_token")
    hs["embeddings.position_embeddings"] = g("pos_embed")
    hs["embeddings.patch_embeddings.projection.weight"] = g("patch_embed.proj.weight")
    hs["embeddings.patch_embeddings.projection.bias"] = g("patch_embed.proj.bias")
    for i in range(12):
        b, h = f"blocks.{i}.", f"encoder.layer.{i}."
        qkv_w, qkv_b = g(b + "attn.qkv.weight"), g(b + "attn.qkv.bias")
        for j, n in enumerate(("query", "key", "value")):
            hs[h + f"attention.attention.{n}.weight"] = qkv_w[j * 768:(j + 1) * 768]
            hs[h + f"attention.attention.{n}.bias"] = qkv_b[j * 768:(j + 1) * 768]
        hs[h + "attention.output.dense.weight"], hs[h + "attention.output.dense.bias"] = g(b + "attn.proj.weight"), g(b + "attn.proj.bias")
        hs[h + "norm1.weight"], hs[h + "norm1.bias"] = g(b + "norm1.weight"), g(b + "norm1.bias")
        hs[h + "norm2.weight"], hs[h + "norm2.bias"] = g(b + "norm2.weight"), g(b + "norm2.bias")
        hs[h + "layer_scale1.lambda1"], hs[h + "layer_scale2.lambda1"] = g(b + "ls1.gamma"), g(b + "ls2.gamma")
        hs[h + "mlp.fc1.weight"], hs[h + "mlp.fc1.bias"] = g(b + "mlp.fc1.weight"), g(b + "mlp.fc1.bias")
        hs[h + "mlp.fc2.weight"], hs[h + "mlp.fc2.bias"] = g(b + "mlp.fc2.weight"), g(b + "mlp.fc2.bias")
    hs["layernorm.weight"], hs["layernorm.bias"] = g("norm.weight"), g("norm.bias")
    m.load_state_dict(hs, strict=True)
    return m


def test_dino_restatement_matches_transformers_at_native_grid():
    """At 518x518 (37x37 patches) no position interpolation happens in either implementation: the block arithmetic
    (LN eps, qkv split, attention scale, LayerScale, GELU MLP, final norm, patch-token selection) must agree."""
    sd = orc.init_state_dict(seed=0, cfg=dict(frames=1))
    g = torch.Generator().manual_seed(5)
    img = torch.rand(1, 3, 518, 518, generator=g)
    mean = torch.tensor(orc._MEAN).view(1, 3, 1, 1)
    std = torch.tensor(orc._STD).view(1, 3, 1, 1)
    with torch.no_grad():
        ours = orc.dino_forward(img, sd)                       # normalises internally
        hf = _hf_model_from(sd)((img - mean) / std).last_hidden_state[:, 1:]
    assert orc.rel_l2(ours, hf) < 2e-5


def test_dino_module_and_functional_agree_at_224():
    """The nn.Module used as the torch.hub stand-in and the functional restatement are the same arithmetic, including the
    bicubic 37x37 -> 16x16 position interpolation with the 0.1 offset."""
    sd = orc.init_state_dict(seed=0, cfg=dict(frames=1))
    m = dino.DinoV2ViTB14().eval()
    m.load_state_dict({k[len("image_encoder.model."):]: v for k, v in sd.items() if k.startswith("image_encoder.model.")}, strict=True)
    g = torch.Generator().manual_seed(6)
    img = torch.rand(2, 3, 224, 224, generator=g)
    mean = torch.tensor(orc._MEAN).view(1, 3, 1, 1)
    std = torch.tensor(orc._STD).view(1, 3, 1, 1)
    with torch.no_grad():
        a = m.forward_features((img - mean) / std)["x_norm_patchtokens"]
        b = orc.dino_forward(img, sd)
    assert a.shape == (2, 256, 768) and orc.rel_l2(a, b) < 1e-5
    pos = dino.interpolate_pos_embed(sd["image_encoder.model.pos_embed"], 16, 16)
    assert pos.shape == (1, 257, 768) and torch.equal(pos[:, 0], sd["image_encoder.model.pos_embed"][:, 0])


def _bicubic_resize_np(table, n_out, coord_scale):
    """Independent bicubic resampler (numpy, no torch): Keys cubic convolution with a = -0.75, half-pixel centres
    (src = (dst + 0.5) * coord_scale - 0.5), border indices clamped -- the published definition of bicubic,
    align_corners=False.  table [M, M, C] -> [n_out, n_out, C]."""
    import numpy as np
    a = -0.75

    def w(t):   # weights of the 4 taps at distances 1+t, t, 1-t, 2-t
        t = np.asarray(t, dtype=np.float64)
        w0 = ((a * (t + 1) - 5 * a) * (t + 1) + 8 * a) * (t + 1) - 4 * a
        w1 = ((a + 2) * t - (a + 3)) * t * t + 1
        w2 = ((a + 2) * (1 - t) - (a + 3)) * (1 - t) * (1 - t) + 1
        w3 = ((a * (2 - t) - 5 * a) * (2 - t) + 8 * a) * (2 - t) - 4 * a
        return np.stack([w0, w1, w2, w3], -1)

    M = table.shape[0]
    src = (np.arange(n_out) + 0.5) * coord_scale - 0.5
    i0 = np.floor(src).astype(np.int64)
    wt = w(src - i0)                                          # [n_out, 4]
    idx = np.clip(i0[:, None] + np.arange(-1, 3)[None], 0, M - 1)   # [n_out, 4]
    t64 = table.astype(np.float64)
    rows = np.einsum("ik,ikjc->ijc", wt, t64[idx])            # resample axis 0 -> [n_out, M, C]
    return np.einsum("jk,ijkc->ijc", wt, rows[:, idx])        # resample axis 1 -> [n_out, n_out, C]


def test_dino_position_interpolation_is_pinned_on_both_upstream_branches():
    """Pins the 37x37 -> 16x16 position-table resize the frozen encoder applies at 224x224 (image_encoder/dinov2.py:55-58 fixes
    the input size, so this is the ONLY grid the hot path ever uses):

    * hub default ``interpolate_offset = 0.1``: F.interpolate(scale_factor=(16.1/37,)*2) maps output pixel centres with the
      GIVEN scale (src = (dst + 0.5) * 37/16.1 - 0.5).  Checked against an independent numpy bicubic (no torch).
    * ``interpolate_offset = 0`` (upstream's other branch = what `transformers` Dinov2 implements, size=(16,16): src scale
      37/16): our restatement with offset=0 equals HF's ``interpolate_pos_encoding`` bit for bit, and the numpy resampler too.

    The two branches differ by design (different sampling coordinates): 16.1 vs 16 in the denominator.  That difference is
    why the end-to-end cross-check against `transformers` runs at the native 518x518 grid and this test pins the rest."""
    import numpy as np
    from transformers import Dinov2Config
    from transformers.models.dinov2.modeling_dinov2 import Dinov2Embeddings
    sd = orc.init_state_dict(seed=0, cfg=dict(frames=1))
    table = sd["image_encoder.model.pos_embed"].float()                       # [1, 1370, 768]
    grid = table[0, 1:].reshape(37, 37, 768).numpy()
    # branch 1: the hub default (what the product folds into its constant table, Motion_Latent_Model._dino_pos)
    ours = dino.interpolate_pos_embed(table, 16, 16)
    ref_np = _bicubic_resize_np(grid, 16, 37.0 / 16.1).reshape(256, 768)
    assert torch.equal(ours[:, 0], table[:, 0])
    assert float(np.abs(ours[0, 1:].numpy() - ref_np).max()) < 5e-6 * float(np.abs(ref_np).max() + 1)
    # branch 2: offset 0 == transformers
    emb = Dinov2Embeddings(Dinov2Config(hidden_size=768, image_size=518, patch_size=14))
    with torch.no_grad():
        emb.position_embeddings.copy_(table)
        hf = emb.interpolate_pos_encoding(torch.zeros(1, 257, 768), 224, 224)
    ours0 = dino.interpolate_pos_embed(table, 16, 16, offset=0)
    assert torch.equal(ours0, hf)
    ref0 = _bicubic_resize_np(grid, 16, 37.0 / 16.0).reshape(256, 768)
    assert float(np.abs(ours0[0, 1:].numpy() - ref0).max()) < 5e-6 * float(np.abs(ref0).max() + 1)
    # and the branches really differ (so the choice matters and is exercised)
    assert orc.rel_l2(ours, ours0) > 1e-3


def test_product_dino_position_table_equals_the_oracle():
    """Motion_Latent_Model._dino_pos (load-time constant fold, CPU-runnable) == the oracle's interpolate_pos_embed."""
    from motion324_b200.model.Pcd_motion import Motion_Latent_Model
    from motion324_b200.utils.config import make_config
    sd = orc.init_state_dict(seed=0, cfg=dict(frames=1))
    model = Motion_Latent_Model(make_config(frames=1))
    table = sd["image_encoder.model.pos_embed"].float()
    got = model._dino_pos(table)
    assert torch.equal(got, dino.interpolate_pos_embed(table, 16, 16).reshape(257, 768))
