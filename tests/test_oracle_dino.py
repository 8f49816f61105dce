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
