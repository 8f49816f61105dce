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
