import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from motion324_b200 import ops
from motion324_b200.model.Pcd_motion import Motion_Latent_Model
from motion324_b200.utils.config import make_config
from motion324_b200.utils import synthetic as orc  # seeded weights / inputs generator
T = 32
model = Motion_Latent_Model(make_config(frames=T)); model.load_state_dict(orc.init_state_dict(0, dict(frames=T))); model = model.to("cuda"); model.eval()
sample = {k: v.to("cuda") for k, v in orc.make_inputs(seed=1, B=1, T=T, N=4096, S=4096).items()}
outs = {}
for pdl_off in (1, 0, 1, 0):
    ops.set_tuning(2, pdl_off)
    for _ in range(3): model(sample)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10): ret = model(sample)
    e.record(); torch.cuda.synchronize()
    outs[pdl_off] = ret.pcd_moved.clone()
    print(f"PDL {'off' if pdl_off else 'on '}: {s.elapsed_time(e)/10:.3f} ms/step  loss {float(ret.loss_metrics.loss):.6f}")
print("bit-identical with/without PDL:", torch.equal(outs[0], outs[1]))
