"""BASELINE.json configs[4]: long-sequence sweep, frames x points, forward+loss frames/s on one GPU (CUDA events)."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from motion324_b200.model.Pcd_motion import Motion_Latent_Model
from motion324_b200.utils.config import make_config
from motion324_b200.utils import synthetic as orc  # seeded weights / inputs generator

rows = []
for T in (8, 32, 128, 256):
    model = Motion_Latent_Model(make_config(frames=T))
    model.load_state_dict(orc.init_state_dict(0, dict(frames=T)))
    model = model.to("cuda"); model.eval()
    for N in (1024, 4096, 16384):
        if T == 256 and N != 4096:
            continue
        sample = {k: v.to("cuda") for k, v in orc.make_inputs(seed=1, B=1, T=T, N=N, S=4096).items()}
        for _ in range(2):
            model(sample)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3 if T >= 128 else 6
        s.record()
        for _ in range(reps):
            ret = model(sample)
        e.record(); torch.cuda.synchronize()
        ms = s.elapsed_time(e) / reps
        mem = torch.cuda.max_memory_allocated() / 2**30
        rows.append(dict(T=T, N=N, ms=ms, fps=T / ms * 1e3, loss=float(ret.loss_metrics.loss), peak_gib=mem))
        print(json.dumps(rows[-1]), flush=True)
    del model
    torch.cuda.empty_cache()
