"""Host-side cost of one forward: time to ENQUEUE all launches (no sync) vs the GPU time of the step."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from motion324_b200.model.Pcd_motion import Motion_Latent_Model
from motion324_b200.utils.config import make_config
from motion324_b200.utils import synthetic as orc  # seeded weights / inputs generator
T = 32
model = Motion_Latent_Model(make_config(frames=T)); model.load_state_dict(orc.init_state_dict(0, dict(frames=T))); model = model.to("cuda"); model.eval()
sample = {k: v.to("cuda") for k, v in orc.make_inputs(seed=1, B=1, T=T, N=4096, S=4096).items()}
for _ in range(3): model(sample)
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter(); s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record(); model(sample); e.record(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"enqueue {1e3*(t1-t0):.2f} ms   gpu {s.elapsed_time(e):.2f} ms   wall {1e3*(t2-t0):.2f} ms")
