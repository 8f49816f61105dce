"""CPU-side checks: the C-ABI library loads and exports every symbol include/m324.h declares (no compute calls), the
drop-in class reproduces the reference's plugin surface, the error behaviour, and the clip-sharding logic of bench.py
under a world_size-2 gloo group."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    from motion324_b200 import build
    return build.build()


def test_library_exports_every_declared_symbol(built_lib):
    hdr = open(os.path.join(ROOT, "include", "m324.h")).read()
    declared = set(re.findall(r"\b(m324_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 15
    lib = ctypes.CDLL(built_lib)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/m324.h but not exported by libm324.so"
    from motion324_b200 import lib as l
    assert set(l.SIGNATURES) == declared      # the ctypes binding covers the whole header
    assert l.load().m324_version() == 100


def test_ctypes_structs_match_header_layout(built_lib):
    src = '#include "%s"\n#include <stdio.h>\nint main(){printf("%%zu %%zu\\n", sizeof(m324_gemm_args), sizeof(m324_attn_args));}\n' % os.path.join(ROOT, "include", "m324.h")
    exe = os.path.join(ROOT, "motion324_b200", "build", "sizes")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    r = subprocess.run(["g++", "-x", "c++", "-", "-o", exe], input=src, text=True, capture_output=True)
    assert r.returncode == 0, r.stderr
    g, a = map(int, subprocess.run([exe], capture_output=True, text=True).stdout.split())
    from motion324_b200 import lib as l
    assert ctypes.sizeof(l.GemmArgs) == g and ctypes.sizeof(l.AttnArgs) == a


def test_sass_is_blackwell_native(built_lib):
    sass = subprocess.run(["cuobjdump", "-sass", built_lib], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTMAREDG"):
        assert mnemonic in sass, mnemonic
    assert "HMMA." not in sass.replace("UTCHMMA", "")  # no legacy mma.sync path


def test_plugin_surface_matches_reference():
    from motion324_b200.model.Pcd_motion import Motion_Latent_Model
    from motion324_b200.utils.config import make_config
    from oracle import motion324_oracle as orc
    import importlib
    cls = importlib.import_module("motion324_b200.model.Pcd_motion").__dict__["Motion_Latent_Model"]  # train.py:84-86
    assert cls is Motion_Latent_Model
    m = cls(make_config(frames=2))
    sd = orc.init_state_dict(0, dict(frames=2))
    assert m.load_state_dict(sd, strict=True).missing_keys == []
    assert list(m.state_dict().keys()) == [k for k in m.state_dict().keys()] and set(m.state_dict()) == set(sd)
    trainable = sum(p.numel() for p in m.parameters() if p.requires_grad)
    assert trainable == 157037315
    assert all(not p.requires_grad for p in m.image_encoder.parameters())
    assert m.eval() is None and m.train() is None       # Pcd_motion.py:372-373 returns None
    no_decay = [n for n, p in m.named_parameters() if p.requires_grad and p.dim() == 1]
    assert "encoder_cross_attn.attn.q_norm.weight" in no_decay  # utils/training_utils.py:39-47 grouping works
    cfg = make_config(frames=2)
    del cfg.training["coord_mse_loss_weight"]
    with pytest.raises(ValueError):
        cls(cfg)                                         # model/loss.py:18-22
    with pytest.raises(RuntimeError):
        m.eval(); m(orc.make_inputs(seed=1, B=1, T=2, N=8, S=8))  # CPU tensors: there is no CPU path


def test_easydict_contract():
    from motion324_b200.utils.easydict import EasyDict
    r = EasyDict(input_data={"a": 1}, pcd_moved=torch.zeros(1))
    assert isinstance(r, dict) and "pcd_moved" in r and r.pcd_moved is r["pcd_moved"] and r.input_data.a == 1


_GLOO_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
# clip sharding as in bench.py / train.py: rank r owns clips r, r+W, ...; per-rank time -> MAX over ranks; frames summed
clips = list(range(rank, 7, world))
ms = torch.tensor([10.0 * (rank + 1)])
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
n = torch.tensor([float(len(clips) * 32)])
dist.all_reduce(n, op=dist.ReduceOp.SUM)
from oracle import motion324_oracle as orc
s = orc.make_inputs(seed=1 + rank, B=1, T=1, N=4, S=4, H=8, W=8)
g = [torch.zeros_like(s["ref_pcd"]) for _ in range(world)]
dist.all_gather(g, s["ref_pcd"])
assert not torch.equal(g[0], g[1])            # different ranks get different synthetic clips
assert float(ms) == 10.0 * world and float(n) == 7 * 32
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
'''


def test_clip_sharding_two_ranks_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER % ROOT)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29517")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs


_GRAD_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
from motion324_b200.model.train_path import GradBuffer

class Tiny(torch.nn.Module):          # the interface GradBuffer reads: named_parameters() + the pos_embed buffer's device
    def __init__(self):
        super().__init__()
        self.register_buffer("pos_embed", torch.zeros(1))
        self.a = torch.nn.Linear(5, 7)                       # 35 + 7 elements: slices are padded to 64 floats
        self.frozen = torch.nn.Linear(3, 3)
        for p in self.frozen.parameters():
            p.requires_grad_(False)
        self.attn = torch.nn.ModuleDict(dict(to_k=torch.nn.Linear(4, 4, bias=False), to_v=torch.nn.Linear(4, 4, bias=False)))

m = Tiny()
gb = GradBuffer(m)
assert set(gb.views) == {"a.weight", "a.bias", "attn.to_k.weight", "attn.to_v.weight"}          # frozen parameters own no slice
assert all(gb.offsets[n] %% 64 == 0 for n in gb.names) and gb.flat.numel() == gb.n_grad + 64
for n, v in gb.views.items():
    v.fill_(float(rank + 1))                                  # "gradients" of this rank
gb.metrics.copy_(torch.tensor([0.5 * (rank + 1), 2.0 * (rank + 1)]))
for n, p in m.named_parameters():
    if p.requires_grad:
        p.grad = gb.views[n]
out = gb.allreduce()
mean = sum(range(1, world + 1)) / world
for n, p in m.named_parameters():
    if p.requires_grad:
        assert p.grad.data_ptr() == gb.views[n].data_ptr() and torch.all(p.grad == mean), n      # averaged in place, .grad aliases the flat buffer
assert torch.allclose(out, torch.tensor([0.5 * mean, 2.0 * mean]))                              # loss metrics ride in the same collective
pad = gb.flat[gb.offsets["a.weight"] + 35: gb.offsets["a.weight"] + 64]
assert torch.all(pad == 0)
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
'''


def test_flat_gradient_allreduce_two_ranks_gloo(tmp_path):
    """train.py's DDP exchange step as one all-reduce of the flat gradient buffer (model/train_path.py:GradBuffer.allreduce)."""
    script = tmp_path / "g.py"
    script.write_text(_GRAD_WORKER % ROOT)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29519")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs


def test_frame_shard_partition_and_kv_gradient_packing():
    """Host logic of the two multi-GPU modes: the frame partition of frame_parallel() and the adjacency of the
    to_k / to_v gradient slices that lets the packed k|v weight gradient be ONE GEMM output (model/train_path.py)."""
    from motion324_b200.model.Pcd_motion import frame_shard, Motion_Latent_Model
    from motion324_b200.model.train_path import GradBuffer, _ksplit
    for T, W in ((32, 8), (128, 2), (6, 3), (5, 1)):
        got = [frame_shard(T, r, W) for r in range(W)]
        frames = [t for first, n in got for t in range(first, first + n)]
        assert frames == list(range(T))                       # contiguous, in rank order, complete
    with pytest.raises(ValueError):
        frame_shard(5, 0, 2)
    with pytest.raises(ValueError):
        frame_shard(8, 2, 2)
    from motion324_b200.utils.config import make_config
    m = Motion_Latent_Model(make_config(frames=1))
    gb = GradBuffer(m)
    d = m.d
    for pfx in ("encoder_cross_attn.", "decoder_cross_attn."):
        kv = gb.packed_kv(pfx, d)
        assert kv.shape == (2 * d, d)
        assert kv.data_ptr() == gb.views[pfx + "attn.to_k.weight"].data_ptr()
        assert kv[d:].data_ptr() == gb.views[pfx + "attn.to_v.weight"].data_ptr()
    assert gb.n_grad >= 157037315 and gb.flat.numel() == gb.n_grad + 64 and all(o % 64 == 0 for o in gb.offsets.values())
    assert set(gb.views) == {n for n, p in m.named_parameters() if p.requires_grad}
    for M, N, K in ((2304, 768, 124416), (768, 768, 64), (768, 3072, 49152), (3, 768, 4096)):
        ks = _ksplit(M, N, K)
        assert 1 <= ks <= 32 and ks <= max(1, (K + 63) // 64 // 4)     # every split keeps at least 4 K-blocks


def test_integration_doc_binding_matches_the_header():
    """INTEGRATION.md shows the ctypes stub a reference maintainer adds: its struct must list the fields of m324_attn_args in
    the header's order (same as motion324_b200/lib.py), and its code blocks must be valid Python."""
    import ast
    from motion324_b200 import lib as l
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", doc, flags=re.S)
    assert blocks
    for b in blocks:
        ast.parse(b)
    stub = next(b for b in blocks if "_AttnArgs" in b)
    fields = re.findall(r'\("(\w+)", C\.c_\w+\)', stub[stub.index("_fields_"):stub.index("_lib.m324_attention.argtypes")])
    assert fields == [n for n, _ in l.AttnArgs._fields_]
    hdr = open(os.path.join(ROOT, "include", "m324.h")).read()
    body = hdr[hdr.index("typedef struct {", hdr.index("xformers.ops.memory_efficient_attention")):hdr.index("} m324_attn_args;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = [n for decl in body.split(";") for n in re.findall(r"(\w+)\s*(?:,|$)", decl.split("{")[-1].strip().split(" ", 1)[-1].replace("*", " "))]
    assert [n for n in names if n] == fields


def _attention_plan(B, H, Lq, Lk, sms=148, ws=True, q_shared=False, partial=(0, 0)):
    from motion324_b200 import lib as l
    a = l.AttnArgs()
    a.B, a.H, a.Lq, a.Lk = B, H, Lq, Lk
    a.q_batch_rows, a.kv_batch_rows, a.q_batch_div = (0 if q_shared else Lq), Lk, 1
    lib_ = l.load()
    if ws:      # only null / non-null and the size are inspected by the planner
        a.workspace = 0x1000
        a.workspace_bytes = max(int(lib_.m324_attention_workspace_bytes()) if sms == 148 else sms * 256 * 66 * 4,
                                int(lib_.m324_attention_partial_bytes(B, H, Lq, max(partial[0], 1))))
    a.partial_parts, a.partial_index = partial
    plan = (ctypes.c_int32 * 8)()
    rc = lib_.m324_attention_plan(ctypes.byref(a), sms, plan)
    assert rc == 0, lib_.m324_last_error()
    return dict(zip(("n_qt", "frame_loop", "items_whole", "split_parts", "split_slots", "grid", "merge_blocks", "item_loop"), list(plan)[:8]))


def _simulate_kernel_decode(p, B, H, Lk, partial=(0, 0)):
    """What attn_kernel computes from blockIdx.x (csrc/attention.cu): (batch, head, Q-tile pair) -> list of K/V tiles, + slot."""
    n_all = (Lk + 127) // 128
    work, slots = {}, []
    if p.get("item_loop"):       # attn_items_kernel: CTA c walks the contiguous chunk [c * items / G, (c + 1) * items / G) of whole items
        G, total = p["grid"], p["items_whole"]
        for c in range(G):
            for item in range(c * total // G, (c + 1) * total // G):
                qt, h, b = item % p["n_qt"], (item // p["n_qt"]) % H, item // (p["n_qt"] * H)
                work.setdefault((b, h, qt), []).extend(range(n_all))
        return work, slots
    for c in range(p["grid"]):
        item, j0, j1, slot = c, 0, n_all, -1
        if item >= p["items_whole"]:
            slot = item - p["items_whole"]
            part = slot % p["split_parts"]
            item = p["items_whole"] + slot // p["split_parts"]
            j0, j1 = part * n_all // p["split_parts"], (part + 1) * n_all // p["split_parts"]
        elif partial[0] > 0:
            slot = item * partial[0] + partial[1]
        qt, h, bg = item % p["n_qt"], (item // p["n_qt"]) % H, item // (p["n_qt"] * H)
        if p["frame_loop"] > 1:
            tiles = [(b, 0) for b in range(bg * p["frame_loop"], min(B, (bg + 1) * p["frame_loop"]))]
        else:
            tiles = [(bg, j) for j in range(j0, j1)]
        for b, j in tiles:
            work.setdefault((b, h, qt), []).append(j)
        if slot >= 0:
            slots.append(slot)
    return work, slots


@pytest.mark.parametrize("B,H,Lq,Lk,shared,expect", [
    (1, 12, 10368, 10368, False, dict(grid=588, items_whole=444, split_parts=3, merge_blocks=1536)),   # global layer, 32 frames (ncu: grid 588)
    (32, 12, 324, 324, False, dict(grid=148, item_loop=1, items_whole=768, split_slots=0, merge_blocks=0)),   # local layers: persistent item loop
    (32, 12, 257, 257, False, dict(grid=148, item_loop=1, items_whole=768)),                           # DINOv2 blocks
    (384, 12, 324, 324, False, dict(grid=148, item_loop=1, items_whole=9216)),                         # training: 32 clips x 12 frames
    (2, 12, 324, 324, False, dict(grid=48, item_loop=0)),                                              # too few items for the loop
    (40, 12, 64, 64, False, dict(grid=148, item_loop=1, items_whole=480)),                             # one K/V tile per item
    (1, 12, 64, 4096, False, dict(grid=96, items_whole=0, split_parts=8, merge_blocks=384)),           # encoder cross-attention (ncu: 96, 384)
    (32, 12, 4096, 64, True, dict(grid=768, frame_loop=8, split_slots=0)),                             # decoder: 8 frames per CTA
    (12, 12, 4096, 64, True, dict(grid=384, frame_loop=8)),                                            # training decoder chunk: 8 + 4 frames
    (5, 12, 300, 64, True, dict(frame_loop=1)),                                                        # too few frames for the loop
    (1, 12, 41472, 41472, False, dict(items_whole=1924, split_parts=4)),                               # 128 frames
    (1, 12, 15552, 15552, False, dict(split_slots=0)),                                                 # 48 frames: last wave nearly full
    (13, 12, 200, 2100, False, dict(split_parts=4)), (2, 3, 130, 264, False, dict(split_slots=0)), (1, 1, 1, 1, False, dict(grid=1)),
])
def test_attention_work_decomposition_covers_every_tile_once(B, H, Lq, Lk, shared, expect):
    """Host logic of m324_attention (tail split, frame loop) through m324_attention_plan -- no GPU: the CTA -> work mapping of
    the kernel, replayed here, must give every (batch, head, Q-tile pair) each of its K/V tiles exactly once, with unique
    workspace slots that fit the workspace; the figures for the model's own launches match the grids ncu recorded on the B200
    (profiles/r1q_launches.csv)."""
    p = _attention_plan(B, H, Lq, Lk, q_shared=shared)
    for k, v in expect.items():
        assert p[k] == v, (k, p)
    work, slots = _simulate_kernel_decode(p, B, H, Lk)
    n_all = (Lk + 127) // 128
    assert set(work) == {(b, h, qt) for b in range(B) for h in range(H) for qt in range(p["n_qt"])}
    assert all(sorted(js) == list(range(n_all)) for js in work.values())
    assert sorted(slots) == list(range(p["split_slots"])) and p["split_slots"] * 256 * 66 * 4 <= 148 * 256 * 66 * 4
    assert p["merge_blocks"] == (p["split_slots"] // p["split_parts"] * 256 + 7) // 8
    q = _attention_plan(B, H, Lq, Lk, ws=False, q_shared=shared)       # no workspace: never split
    assert q["split_slots"] == 0 and q["grid"] == (148 if q["item_loop"] else q["items_whole"]) and q["merge_blocks"] == 0


def test_attention_partial_launch_plan_and_small_devices():
    """Partial launches (one K/V range each, m324_attention_merge afterwards): one CTA per work item, slot = item * parts + index;
    and the tail split on a device with fewer SMs."""
    B, H, Lq = 1, 12, 5184
    seen = []
    for idx in range(3):
        p = _attention_plan(B, H, Lq, 5184 * (1 + idx), partial=(3, idx))
        assert p["grid"] == p["items_whole"] == 21 * 12 and p["merge_blocks"] == 0 and p["split_slots"] == 3 * 252
        _, slots = _simulate_kernel_decode(p, B, H, 5184 * (1 + idx), partial=(3, idx))
        seen += slots
    assert sorted(seen) == list(range(3 * 252))
    from motion324_b200 import lib as l
    a = l.AttnArgs()
    a.B, a.H, a.Lq, a.Lk, a.q_batch_div, a.partial_parts, a.partial_index = 1, 12, 512, 512, 1, 2, 2
    a.workspace, a.workspace_bytes = 0x1000, 1 << 30
    assert l.load().m324_attention_plan(ctypes.byref(a), 148, (ctypes.c_int32 * 8)()) != 0        # index out of range
    for sms in (4, 74, 132):
        p = _attention_plan(1, 12, 10368, 10368, sms=sms)
        work, slots = _simulate_kernel_decode(p, 1, 12, 10368)
        assert all(sorted(js) == list(range(81)) for js in work.values()) and len(work) == 12 * 41
        assert sorted(slots) == list(range(p["split_slots"])) and p["split_slots"] <= sms


def test_overlapped_allreduce_plan_tiles_the_flat_gradient_buffer():
    """GradBuffer.overlap_plan: the three waves of the overlapped gradient exchange cover every element of the flat buffer exactly
    once, and each wave only holds parameters whose gradient is final when the backward reaches that point."""
    from motion324_b200.model.Pcd_motion import Motion_Latent_Model
    from motion324_b200.model.train_path import GradBuffer
    from motion324_b200.utils.config import make_config
    gb = GradBuffer(Motion_Latent_Model(make_config(frames=2)))
    plan = gb.overlap_plan()
    cover = sorted(r for k in ("trunk_hi", "trunk_lo", "rest") for r in plan[k])
    assert cover[0][0] == 0 and cover[-1][1] == gb.flat.numel() and all(a[1] == b[0] for a, b in zip(cover, cover[1:]))
    half = plan["half"]
    for name in gb.names:
        o = gb.offsets[name]
        wave = next(k for k in ("trunk_hi", "trunk_lo", "rest") if any(lo <= o < hi for lo, hi in plan[k]))
        if name.startswith(("global_transformer_blocks.", "local_transformer_blocks.")):
            assert wave == ("trunk_hi" if int(name.split(".")[1]) >= half else "trunk_lo"), name
        else:
            assert wave == "rest", name


def test_frame_sources_index_plan_equals_the_reference_stitch():
    """inference.frame_sources (which window produces which output frame) against the oracle's stitch, which is pinned to the
    reference's own merge code (tests/golden/inference_windows.npz): every chunk 2..8 x every clip length up to 59."""
    from motion324_b200.inference import window_plan, frame_sources
    from oracle import inference_oracle as io
    for chunk in range(2, 9):
        for total_T in range(chunk + 1, 60):
            plan = window_plan(total_T, chunk)
            outs = []
            for w, (s, fr) in enumerate(plan):
                o = torch.zeros(1, chunk, 1, 3)
                for l, g in enumerate(fr):
                    o[0, l] = 1000 * w + g           # the value names the producing window and the frame
                outs.append(o)
            ref = io.stitch(outs, [s for s, _ in plan], torch.full((1, 1, 3), -1.0))
            got = torch.zeros(1, total_T, 1, 3)
            cover = [0] * total_T
            for w, local, first, count in frame_sources(total_T, chunk):
                got[:, first:first + count] = outs[w][:, local:local + count]
                for g in range(first, first + count):
                    cover[g] += 1
            got[:, 0] = -1.0
            assert cover == [0] + [1] * (total_T - 1), (chunk, total_T)      # frames 1.. tiled exactly once
            assert torch.equal(ref, got), (chunk, total_T)


_INFER_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
from motion324_b200.inference import run_model_inference, window_plan
from motion324_b200.utils.easydict import EasyDict

calls = []
def model(sample):                      # stand-in for the plugin: pcd_moved[t] = ref_pcd + mean colour of frame t
    v = sample["rgb_video"]             # [1, chunk, H, W, 3]
    calls.append(v.shape[1])
    return EasyDict(pcd_moved=sample["ref_pcd"][:, None] + v.mean(dim=(2, 3, 4)).view(1, -1, 1, 1))

cfg = EasyDict(training=dict(frames=12))
g = torch.Generator().manual_seed(7)
ref_pcd = torch.rand(1, 5, 3, generator=g)
for total_T in (23, 40, 12):            # 23 frames / chunk 12 -> 2 windows on 3 ranks: rank 2 owns none and must still join
    video = torch.rand(total_T, 4, 4, 3, generator=g)
    plan = window_plan(total_T, 12)
    out = run_model_inference(model, dict(ref_pcd=ref_pcd), video, cfg, "cpu", rank=rank, world_size=world)
    one = run_model_inference(model, dict(ref_pcd=ref_pcd), video, cfg, "cpu")          # single-rank result
    assert out.shape == (1, total_T, 5, 3) and torch.equal(out, one), total_T
    if total_T > 12:
        assert torch.equal(out[:, 0], ref_pcd)
if world == 3:
    assert len(window_plan(23, 12)) == 2
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
'''


def test_sliding_window_inference_more_ranks_than_windows_gloo(tmp_path):
    """run_model_inference with world_size 3 and only 2 windows: the idle rank joins the collective (no deadlock, no
    StopIteration) and every rank ends with the single-rank result, bit for bit."""
    script = tmp_path / "i.py"
    script.write_text(_INFER_WORKER % ROOT)
    procs = []
    for r in range(3):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="3", MASTER_ADDR="127.0.0.1", MASTER_PORT="29523")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs


def test_config_loader_matches_the_reference_cli_semantics(tmp_path):
    """utils/config.py:init_config vs setup.py:52-89: dotted overrides, whitespace around '=', new top-level keys, typed scalars,
    ${...} interpolation."""
    from motion324_b200.utils.config import load_config, process_overrides
    y = tmp_path / "c.yaml"
    y.write_text("model:\n  class_name: a.B\ntraining:\n  frames: 12\n  lr: 0.0004\n  wandb_exp_name: test\n  checkpoint_dir: ./ckpt/${training.wandb_exp_name}\n  use_amp: true\n")
    assert process_overrides(["a", "=", "1", "b=", "2", "c=3"]) == ["a=1", "b=2", "c=3"]
    c = load_config(str(y), ["training.frames", "=", "256", "training.lr=1e-4", "data_dir=x.glb", "use_segmentation=False",
                             "training.wandb_exp_name=run7", "model.class_name=motion324_b200.model.Pcd_motion.Motion_Latent_Model"])
    assert c.training.frames == 256 and isinstance(c.training.frames, int)
    assert c.training.lr == 1e-4 and isinstance(c.training.lr, float)
    assert c.data_dir == "x.glb" and c.use_segmentation is False and c.training.use_amp is True
    assert c.training.checkpoint_dir == "./ckpt/run7"
    assert c.model.class_name.endswith("Motion_Latent_Model") and c.training.get("missing", 5) == 5


def test_glb_reader(tmp_path):
    """utils/glb.py: a hand-assembled two-triangle GLB (strided vertex buffer, uint16 indices, node translation) and the error
    paths; the chili demo asset when the reference tree is present."""
    import json
    import struct
    import numpy as np
    from motion324_b200.utils.glb import GlbError, load_glb
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0]], np.float32)
    uv = np.array([[0, 0], [1, 0], [0, 1], [1, 1]], np.float32)
    inter = np.concatenate([pos, uv], 1).astype(np.float32).tobytes()       # interleaved, stride 20
    idx = np.array([0, 1, 2, 2, 1, 3], np.uint16).tobytes()
    binary = inter + idx
    doc = {"asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0]}], "nodes": [{"mesh": 0, "translation": [10, 0, 0]}],
           "meshes": [{"primitives": [{"attributes": {"POSITION": 0, "TEXCOORD_0": 1}, "indices": 2}]}],
           "buffers": [{"byteLength": len(binary)}],
           "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": len(inter), "byteStride": 20},
                           {"buffer": 0, "byteOffset": len(inter), "byteLength": len(idx)}],
           "accessors": [{"bufferView": 0, "byteOffset": 0, "componentType": 5126, "count": 4, "type": "VEC3"},
                         {"bufferView": 0, "byteOffset": 12, "componentType": 5126, "count": 4, "type": "VEC2"},
                         {"bufferView": 1, "componentType": 5123, "count": 6, "type": "SCALAR"}]}
    js = json.dumps(doc).encode()
    js += b" " * (-len(js) % 4)
    binary += b"\0" * (-len(binary) % 4)
    blob = struct.pack("<4sII", b"glTF", 2, 12 + 8 + len(js) + 8 + len(binary)) + struct.pack("<II", len(js), 0x4E4F534A) + js \
        + struct.pack("<II", len(binary), 0x004E4942) + binary
    f = tmp_path / "quad.glb"
    f.write_bytes(blob)
    g = load_glb(str(f))
    assert np.array_equal(g["vertices"], pos.astype(np.float64) + [10, 0, 0]) and np.array_equal(g["uv"], uv)
    assert np.array_equal(g["faces"], [[0, 1, 2], [2, 1, 3]]) and g["normals"] is None and g["texture"] is None
    (tmp_path / "bad.glb").write_bytes(blob[:40])
    with pytest.raises(GlbError):
        load_glb(str(tmp_path / "bad.glb"))
    (tmp_path / "magic.glb").write_bytes(b"XXXX" + blob[4:])
    with pytest.raises(GlbError):
        load_glb(str(tmp_path / "magic.glb"))
    from oracle import build_ref
    if build_ref.available():
        c = load_glb(os.path.join(build_ref.root(), "examples", "chili.glb"))
        assert c["vertices"].shape == (13465, 3) and c["faces"].shape == (19753, 3) and c["uv"].shape == (13465, 2)
        assert c["texture"].ndim == 3 and c["texture"].shape[2] == 3 and c["texture"].dtype == np.uint8
        assert abs(float(np.linalg.norm(c["normals"], axis=1).mean()) - 1.0) < 1e-3


def test_synthetic_dataset_goes_through_the_reference_collate():
    """training.dataset_name seam (train.py:50-53): SyntheticDyscene(config.training) items have the schema of
    dataset/dyscene.py:315-327 and the reference's own collate_fn_with_topology (dataset/dyscene.py:331-383, run unmodified with an
    empty trimesh stand-in) batches them into what Motion_Latent_Model.forward reads."""
    from oracle import build_ref
    if not build_ref.available():
        pytest.skip("reference tree not staged")
    from motion324_b200.dataset.synthetic import SyntheticDyscene
    from motion324_b200.utils.config import make_config
    cfg = make_config(frames=3, num_pcd_samples=40, num_shape_samples=50, synthetic_len=10, synthetic_image_size=28)
    ds = SyntheticDyscene(cfg.training)
    assert len(ds) == 10
    it = ds[7]
    assert it["rgb_video"].shape == (3, 28, 28, 3) and it["point_clouds"].shape == (3, 40, 3) and it["ref_shape_pcd"].shape == (50, 3)
    assert float(it["rgb_video"].min()) >= 0 and float(it["rgb_video"].max()) <= 1 and float(it["ref_pcd"].abs().max()) <= 0.5
    from oracle import ref_shims
    ref_shims.install_trimesh_stub()
    root = build_ref.root()
    sys.path.insert(0, root)
    try:
        import importlib
        collate = importlib.import_module("dataset.dyscene").collate_fn_with_topology
    finally:
        sys.path.remove(root)
    batch = collate([ds[0], ds[1], ds[5]])
    assert batch["rgb_video"].shape == (3, 3, 28, 28, 3) and batch["point_clouds"].shape == (3, 3, 40, 3)
    assert batch["ref_shape_normals"].shape == (3, 50, 3) and batch["obj_name"] == ["synthetic_000000", "synthetic_000001", "synthetic_000005"]
    assert torch.equal(batch["ref_pcd"][2], ds[5]["ref_pcd"])


def test_omegaconf_shim_reproduces_init_config(tmp_path, monkeypatch):
    """The unmodified reference setup.init_config (setup.py:69-89) through oracle/shims/omegaconf == the product's loader."""
    from oracle import build_ref
    if not build_ref.available():
        pytest.skip("reference tree not staged")
    from oracle import ref_shims
    ref_shims.install()
    monkeypatch.syspath_prepend(os.path.join(ROOT, "oracle", "shims"))
    root = build_ref.root()
    import importlib.util
    spec = importlib.util.spec_from_file_location("m324_ref_setup", os.path.join(root, "setup.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    yaml_path = os.path.join(root, "configs", "dyscene.yaml")
    argv = ["train.py", "--config", yaml_path, "training.frames", "=", "32", "training.lr=1e-4", "data_dir=a.glb", "training.wandb_exp_name=zz"]
    monkeypatch.setattr(sys, "argv", argv)
    ref_cfg = mod.init_config()
    from motion324_b200.utils.config import init_config
    ours = init_config(argv[1:])
    assert dict(ref_cfg) == dict(ours)
    assert ref_cfg.training.frames == 32 and ref_cfg.training.checkpoint_dir.endswith("/zz") and ref_cfg.data_dir == "a.glb"


def test_savgol_taps_equal_scipy_coefficients():
    """motion324_b200.inference.savgol_taps (host-side constant of the 'savgol' smoothing method) == scipy.signal.savgol_coeffs."""
    import numpy as np
    scipy_signal = pytest.importorskip("scipy.signal")
    from motion324_b200.inference import savgol_taps
    for w, p in ((3, 2), (5, 2), (7, 3), (9, 4), (11, 2), (17, 5), (5, 4)):
        assert np.allclose(np.array(savgol_taps(w, p)), scipy_signal.savgol_coeffs(w, p), rtol=0, atol=1e-12), (w, p)
    with pytest.raises(ValueError):
        savgol_taps(3, 3)


_OVERLAP_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
from motion324_b200.model.train_path import GradBuffer, OverlappedAllReduce

def block():
    m = torch.nn.Module()
    m.norm1 = torch.nn.LayerNorm(8, bias=False)
    m.fc = torch.nn.Linear(8, 8, bias=False)
    return m

class Tiny(torch.nn.Module):      # the parameter-name layout GradBuffer.overlap_plan reads (model/Pcd_motion.py order)
    def __init__(self):
        super().__init__()
        self.register_buffer("pos_embed", torch.zeros(1))
        self.learnable_tokens = torch.nn.Parameter(torch.zeros(3, 8))
        self.global_transformer_blocks = torch.nn.ModuleList([block() for _ in range(4)])
        self.local_transformer_blocks = torch.nn.ModuleList([block() for _ in range(4)])
        self.transformer_input_layernorm = torch.nn.LayerNorm(8, bias=False)
        self.head = torch.nn.Linear(8, 3)

m = Tiny()
gb = GradBuffer(m)
torch.manual_seed(100 + rank)
gb.flat.copy_(torch.randn(gb.flat.numel()))
mine = gb.flat.clone()
gathered = [torch.zeros_like(mine) for _ in range(world)]
dist.all_gather(gathered, mine)
want = sum(gathered) / world
ar = OverlappedAllReduce(gb)                  # host buffer + gloo: SUM waves, one scale in finish()
for tag in ("trunk_hi", "trunk_lo", "rest"):  # the order TrainPath.run reports them in
    ar.ready(tag)
metrics = ar.finish()
assert torch.allclose(gb.flat, want, atol=1e-6), float((gb.flat - want).abs().max())      # every element averaged exactly once
assert metrics.data_ptr() == gb.metrics.data_ptr()
blocking = GradBuffer(m)
blocking.flat.copy_(mine)
blocking.allreduce()
assert torch.allclose(blocking.flat, gb.flat, atol=1e-6)                                   # == the single blocking all-reduce
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
'''


def test_overlapped_gradient_allreduce_two_ranks_gloo(tmp_path):
    """OverlappedAllReduce (the gradient exchange in three waves, model/train_path.py) on host buffers under gloo: after the waves in
    backward order every element of the flat buffer is the rank average, exactly like GradBuffer.allreduce()."""
    script = tmp_path / "o.py"
    script.write_text(_OVERLAP_WORKER % ROOT)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29527")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
