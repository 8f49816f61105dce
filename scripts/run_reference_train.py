"""Run the UNMODIFIED reference train.py (staged under oracle/_ref, or /root/reference in the build container) against the drop-in
model class and the synthetic dataset -- BASELINE config (c):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/run_reference_train.py --steps 6 [--model reference] [key=value ...]

Nothing of train.py is edited: the modules this image lacks are provided as shims before it starts (omegaconf -> oracle/shims,
easydict, an empty trimesh for ``from dataset.dyscene import collate_fn_with_topology``; xformers / torch.hub only when the
reference's own model class is selected), wandb runs in disabled mode with a dummy key file (configs/api_keys.yaml is empty in
the reference), and the overrides select ``model.class_name`` / ``training.dataset_name`` through the reference's
own dotted-path seams (train.py:50-53, 84-86).  Rank 0 prints train.py's per-step log; the last line is a JSON summary."""
import argparse
import io
import json
import os
import re
import runpy
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--model", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--keep-checkpoint", action="store_true")
    ap.add_argument("overrides", nargs="*")
    args = ap.parse_args()
    from oracle import build_ref, ref_shims
    ref_root = build_ref.root()
    sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
    ref_shims.install(attention="flash" if args.model == "reference" else "exact")      # easydict (+ xformers / hub for the reference class)
    ref_shims.install_trimesh_stub()
    os.environ.setdefault("WANDB_MODE", "disabled")
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    rank = int(os.environ.get("RANK", "0"))
    run_dir = os.environ.get("M324_TRAIN_RUN_DIR") or tempfile.mkdtemp(prefix="m324_trainpy_")
    os.makedirs(run_dir, exist_ok=True)
    key_file = os.path.join(run_dir, f"api_keys_{rank}.yaml")
    with open(key_file, "w") as f:
        f.write("wandb: disabled-dummy-key\n")
    cls = "motion324_b200.model.Pcd_motion.Motion_Latent_Model" if args.model == "ours" else "model.Pcd_motion.Motion_Latent_Model"
    overrides = [f"model.class_name={cls}", "training.dataset_name=motion324_b200.dataset.synthetic.SyntheticDyscene",
                 f"training.batch_size_per_gpu={args.batch}", f"training.train_steps={args.steps}", f"training.stop_steps={args.steps}",
                 "training.warmup=2", "training.num_workers=4", f"training.api_key_path={key_file}",
                 f"training.checkpoint_dir={os.path.join(run_dir, 'ckpt')}", "training.checkpoint_every=1000000",
                 "training.print_every=1"] + args.overrides
    if args.model == "reference":
        # the randomly initialised stand-in ViT gives the reference class a first-step gradient norm of ~14 > 5 x grad_clip_norm, and
        # train.py:198-201 then skips every optimizer step, so its while-loop (train.py:135) never reaches train_steps: lift the guard
        overrides += ["model.video_encoder.transformer.drop_rate=0.1", "training.allowed_gradnorm_factor=1000000"]
    sys.argv = ["train.py", "--config", os.path.join(ref_root, "configs", "dyscene.yaml")] + overrides
    if ref_root not in sys.path:
        sys.path.insert(0, ref_root)
    os.chdir(run_dir)
    log = io.StringIO()

    class Tee:
        def __init__(self, a, b):
            self.a, self.b = a, b

        def write(self, s):
            self.a.write(s)
            self.b.write(s)

        def flush(self):
            self.a.flush()

        def isatty(self):
            return False

    out0 = sys.stdout
    sys.stdout = Tee(out0, log)
    try:
        runpy.run_path(os.path.join(ref_root, "train.py"), run_name="__main__")
    finally:
        sys.stdout = out0
    if rank == 0:
        text = log.getvalue()
        times = [float(x) for x in re.findall(r"Iter Time: ([0-9.]+)s", text)]
        losses = [float(x) for x in re.findall(r"\bloss: ([0-9.]+)", text)]
        world = int(os.environ.get("WORLD_SIZE", "1"))
        steady = sorted(times[2:])[: max(1, len(times[2:]))] if len(times) > 2 else times
        med = steady[len(steady) // 2] if steady else None
        ckpts = [f for f in os.listdir(os.path.join(run_dir, "ckpt"))] if os.path.isdir(os.path.join(run_dir, "ckpt")) else []
        print(json.dumps({"what": "unmodified reference train.py (oracle/_ref) + " + cls + " + SyntheticDyscene", "world_size": world,
                          "batch_size_per_gpu": args.batch, "steps_logged": len(times), "iter_time_s": times, "loss": losses,
                          "median_iter_time_s_after_2": med,
                          "frames_per_s": (world * args.batch * 12 / med) if med else None,
                          "checkpoint_files": sorted(f for f in ckpts if f.endswith(".pt"))}))
        if not args.keep_checkpoint:
            import shutil
            shutil.rmtree(os.path.join(run_dir, "ckpt"), ignore_errors=True)


if __name__ == "__main__":
    main()
