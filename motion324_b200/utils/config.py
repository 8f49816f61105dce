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
