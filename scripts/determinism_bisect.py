"""Which stage first differs between synced and back-to-back forwards (workspace snapshots cloned on the main stream after each forward)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motion324_b200.model.Pcd_motion import Motion_Latent_Model
from motion324_b200.utils.config import make_config
from motion324_b200.utils import synthetic as syn
T, N = 32, 4096
model = Motion_Latent_Model(make_config(frames=T)); model.load_state_dict(syn.init_state_dict(0, dict(frames=T)), strict=True)
model = model.to("cuda"); model.eval()
if os.environ.get("M324_NO_SIDE"):
    model._side_stream = lambda: torch.cuda.current_stream()
sample = {k: v.to("cuda") for k, v in syn.make_inputs(seed=1, B=1, T=T, N=N, S=N).items()}
names = ["shape_feat", "enc_kv16", "enc_o16", "mesh_feat", "patch_embed", "dino_x", "trunk_x", "dec_feat", "dec_q16", "dec_kv16", "dec_x", "dec_hpart"]
def snap(r):
    ws = {k[0]: v for k, v in model._ws.items()}
    d = {n: ws[n].clone() for n in names if n in ws}
    d["out"] = r.pcd_moved.clone()
    return d
r = model(sample); torch.cuda.synchronize()
ref = snap(r); torch.cuda.synchronize()
for burst in range(4):
    snaps = []
    for i in range(10):
        snaps.append(snap(model(sample)))
    torch.cuda.synchronize()
    for i, s in enumerate(snaps):
        bad = [n for n in list(ref) if not torch.equal(s[n], ref[n])]
        if bad:
            print(f"burst {burst} run {i}: differing stages (in pipeline order): {bad}")
            break
    else:
        print(f"burst {burst}: all identical")
