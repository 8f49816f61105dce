"""Bitwise determinism of the benchmark forward: runs separated by a device sync and bursts of back-to-back runs (no sync, PDL overlap
between consecutive kernels and consecutive forwards) must all give the same pcd_moved / loss bits.  Knobs: 2 = PDL off, 5 = LDG residual path."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motion324_b200 import ops
from motion324_b200.model.Pcd_motion import Motion_Latent_Model
from motion324_b200.utils.config import make_config
from motion324_b200.utils import synthetic as syn
T, N = 32, 4096
model = Motion_Latent_Model(make_config(frames=T)); model.load_state_dict(syn.init_state_dict(0, dict(frames=T)), strict=True)
model = model.to("cuda"); model.eval()
sample = {k: v.to("cuda") for k, v in syn.make_inputs(seed=1, B=1, T=T, N=N, S=N).items()}
r = model(sample); torch.cuda.synchronize()
ref = r.pcd_moved.clone(); ref_loss = r.loss_metrics.loss.clone()
bad = 0
for pdl_off in (0, 1):
    for k5 in (0, 1):
        ops.set_tuning(2, pdl_off); ops.set_tuning(5, k5)
        for burst in range(6):
            outs = []
            for i in range(12):
                r = model(sample)
                outs.append((r.pcd_moved.clone(), r.loss_metrics.loss.clone()))
            torch.cuda.synchronize()
            diff = [i for i, (o, l) in enumerate(outs) if not (torch.equal(o, ref) and torch.equal(l, ref_loss))]
            if diff:
                bad += 1
                o, l = outs[diff[0]]
                nd = int((o != ref).sum())
                fr = sorted(set((o != ref).nonzero()[:, 1].tolist()))[:8]
                print(f"pdl_off={pdl_off} knob5={k5} burst {burst}: runs {diff} differ; first: {nd} elements, frames {fr}, max abs {float((o - ref).abs().max()):.3e}, loss {float(l):.10f} vs {float(ref_loss):.10f}")
print("mismatching bursts:", bad)
