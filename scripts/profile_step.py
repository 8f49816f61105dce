"""One profiled forward+loss step at the bench workload (32 frames x 4096 points), bracketed by
cudaProfilerStart/Stop so that `ncu --profile-from-start off` sees exactly one step.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python scripts/profile_step.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from motion324_b200.model.Pcd_motion import Motion_Latent_Model  # noqa: E402
from motion324_b200.utils.config import make_config  # noqa: E402
from motion324_b200.utils import synthetic as orc  # seeded weights / inputs generator

T = int(os.environ.get("M324_T", "32"))
N = int(os.environ.get("M324_N", "4096"))
model = Motion_Latent_Model(make_config(frames=T))
model.load_state_dict(orc.init_state_dict(0, dict(frames=T)), strict=True)
model = model.to("cuda")
model.eval()
sample = {k: v.to("cuda") for k, v in orc.make_inputs(seed=1, B=1, T=T, N=N, S=N).items()}
for _ in range(2):
    model(sample)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ret = model(sample)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", float(ret.loss_metrics.loss))
