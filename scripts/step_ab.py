"""A/B of tuning knobs inside ONE process on the benchmark step (config (b)): alternating rounds, CUDA events.
    python scripts/step_ab.py 4:0,1 1:0,2      # knob:value,value ..."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from motion324_b200 import ops
from motion324_b200.model.Pcd_motion import Motion_Latent_Model
from motion324_b200.utils.config import make_config
from motion324_b200.utils import synthetic as syn

T, N = int(os.environ.get("M324_T", "32")), int(os.environ.get("M324_N", "4096"))
model = Motion_Latent_Model(make_config(frames=T))
model.load_state_dict(syn.init_state_dict(0, dict(frames=T)), strict=True)
model = model.to("cuda"); model.eval()
sample = {k: v.to("cuda") for k, v in syn.make_inputs(seed=1, B=1, T=T, N=N, S=N).items()}


def run(steps=10):
    for _ in range(3):
        model(sample)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        r = model(sample)
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / steps, float(r.loss_metrics.loss)


for spec in sys.argv[1:]:
    if spec.startswith("decode_rows:"):      # Motion_Latent_Model.max_decode_rows: rows of one decoder chunk (frames x points)
        vals = [int(v) for v in spec.split(":")[1].split(",")]
        res = {v: [] for v in vals}
        for rnd in range(3):
            for v in vals:
                model.max_decode_rows = v
                res[v].append(run()[0])
        model.max_decode_rows = 1 << 18
        print(json.dumps({"max_decode_rows": {str(v): [round(x, 3) for x in xs] for v, xs in res.items()}}), flush=True)
        continue
    knob, vals = spec.split(":")
    vals = [int(v) for v in vals.split(",")]
    res = {v: [] for v in vals}
    for rnd in range(4):
        for v in vals:
            ops.set_tuning(int(knob), v)
            ms, loss = run()
            res[v].append(ms)
    ops.set_tuning(int(knob), vals[0])
    print(json.dumps({"knob": int(knob), "ms_per_step": {str(v): [round(x, 3) for x in xs] for v, xs in res.items()},
                      "best": {str(v): min(xs) for v, xs in res.items()}, "loss": loss}), flush=True)
