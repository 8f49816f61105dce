"""Generate the golden fixtures that pin oracle/ against the UNMODIFIED reference.

Runs only in the build container (needs /root/reference, imported through oracle/ref_shims.py).
For each case: weights = oracle.init_state_dict(seed) loaded into the reference class with
strict=True (proves the state_dict key/shape layout), inputs = oracle.make_inputs(seed); the
reference forward (fp32, eval, no autocast) produces pcd_moved / loss, stored with a few stage
tensors.  Weights and inputs are NOT stored: they are regenerated from the seeds.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import motion324_oracle as orc  # noqa: E402
from oracle import ref_shims  # noqa: E402

CASES = {
    # name: (model frames, input T, N, S, H, W)
    "a_T1_N512": dict(frames=1, T=1, N=512, S=512, H=224, W=224),          # BASELINE config (a)
    "resize_T3_N300": dict(frames=4, T=3, N=300, S=700, H=224, W=224),     # pos-embed trilinear resize path
    "chunk_T2_N4200": dict(frames=2, T=2, N=4200, S=256, H=160, W=192),    # eval N-chunking + bilinear resize
    "b_T32_N4096": dict(frames=32, T=32, N=4096, S=4096, H=224, W=224),    # BASELINE config (b): the benchmarked workload
}


POINT_STRIDE = {"b_T32_N4096": 4}   # the config-(b) fixture keeps every 4th point of every frame (393 KB) + fp64 checksums of all


def _stored(name, pcd):
    return np.ascontiguousarray(pcd[:, :, ::POINT_STRIDE.get(name, 1)])


def main():
    torch.set_num_threads(os.cpu_count())
    out_dir = os.path.dirname(os.path.abspath(__file__))
    only = sys.argv[1:]
    for name, c in CASES.items():
        if only and name not in only:
            continue
        cfg = dict(frames=c["frames"])
        sd = orc.init_state_dict(seed=0, cfg=cfg)
        model = ref_shims.build_reference_model(frames=c["frames"])
        missing = model.load_state_dict(sd, strict=True)
        assert not missing.missing_keys and not missing.unexpected_keys
        ref_sd = model.state_dict()
        assert list(ref_sd.keys()) == list(sd.keys()) or set(ref_sd.keys()) == set(sd.keys())
        for k, v in ref_sd.items():
            assert tuple(v.shape) == tuple(sd[k].shape), k
        n_train = sum(p.numel() for p in model.parameters() if p.requires_grad)
        sample = orc.make_inputs(seed=1, B=1, T=c["T"], N=c["N"], S=c["S"], H=c["H"], W=c["W"])
        with torch.no_grad():
            ret = model(dict(sample))
        assert isinstance(ret, dict) and "pcd_moved" in ret
        np.savez_compressed(
            os.path.join(out_dir, f"{name}.npz"),
            pcd_moved=_stored(name, ret["pcd_moved"].numpy()), pcd_moved_sum=np.float64(ret["pcd_moved"].double().sum().item()),
            pcd_moved_sqsum=np.float64(ret["pcd_moved"].double().pow(2).sum().item()), point_stride=np.int64(POINT_STRIDE.get(name, 1)), loss=ret["loss_metrics"]["loss"].numpy(),
            xyz_loss=ret["loss_metrics"]["xyz_loss"].numpy(), n_trainable=np.int64(n_train),
            keys=np.array(sorted(ref_sd.keys())),
        )
        print(name, tuple(ret["pcd_moved"].shape), float(ret["loss_metrics"]["loss"]), "trainable", n_train)


if __name__ == "__main__":
    main()
