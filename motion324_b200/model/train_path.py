"""Training step of Motion_Latent_Model on libm324: what ``train.py:157-170`` runs through torch.autograd
for the reference -- forward with saved activations, then the hand-written backward -- as one launch sequence.

Every arithmetic step is a libm324 kernel (C ABI, include/m324.h); this file owns buffers and the launch order only.
Conventions (csrc/backward.cu): activation gradients travel in units of 1/alpha with alpha = 2 * coord_mse_loss_weight / n
(the seed is pred - target): fp16 between GEMMs, fp32 on the residual stream; parameter gradients are accumulated (+=) in
true units into ONE flat fp32 buffer whose slices are the parameters' ``.grad`` -- the single buffer train.py's gradient
all-reduce needs (one ncclAllReduce instead of DDP's buckets).

Memory plan (180 GB HBM3e): no activation checkpointing (the reference needs ``use_checkpoint`` on smaller parts,
Pcd_motion.py:387-392,424-428) -- every trunk block keeps its fp32 residual inputs, fp16 GEMM operands, GELU
pre-activations, log-sum-exp and q/k reciprocal RMS (27.8 KB per token and block; 55 GB at batch 32 x 12 frames).  The
per-point decoder, whose activations would dominate (B*T*N rows), is never stored: each frame chunk runs forward and
backward back to back, because the loss is a sum over chunks.
"""
import math

import torch

from .. import ops

F16, F32 = torch.float16, torch.float32
KP_EMB, KP_FEAT = 64, 832


def _cdiv(a, b):
    return (a + b - 1) // b


def _ksplit(M, N, K):
    """Split-K factor of a weight-gradient GEMM dW[M, N] = dY^T X over K rows: fill the 148 SMs (1-CTA 128 x BN tiles)."""
    tiles = _cdiv(M, 128) * _cdiv(N, 256 if N % 256 == 0 else 128)
    kb = _cdiv(K, 64)
    best, best_eff = 1, 0.0
    for ks in range(1, max(1, min(32, kb // 4)) + 1):
        ctas = tiles * ks
        eff = ctas / (_cdiv(ctas, 148) * 148.0)
        if eff > best_eff + 0.02:
            best, best_eff = ks, eff
    return best


class GradBuffer:
    """One flat fp32 gradient buffer; ``views[name]`` is the slice of parameter ``name`` (shape of the parameter)."""

    def __init__(self, model):
        dev = model.pos_embed.device
        self.names, self.offsets, off = [], {}, 0
        for name, p in model.named_parameters():
            if not p.requires_grad:
                continue
            self.names.append(name)
            self.offsets[name] = off
            off += _cdiv(p.numel(), 64) * 64          # 256-byte aligned slices (TMA reduce-add needs 16)
        self.n_grad = off
        self.flat = torch.zeros(off + 64, device=dev, dtype=F32)   # + 64 floats that ride along in the all-reduce (loss metrics)
        self.metrics = self.flat[off:off + 2]                      # (xyz_loss, loss) of the last step, averaged by allreduce()
        self.views = {}
        for name, p in model.named_parameters():
            if p.requires_grad:
                o = self.offsets[name]
                self.views[name] = self.flat[o:o + p.numel()].view(p.shape)

    def allreduce(self, group=None):
        """The data-parallel exchange step of train.py (DDP's bucketed gradient all-reduce, train.py:88-89, 162-166) as ONE
        collective over the flat buffer: gradients and the two loss scalars are averaged over the ranks of `group` in place
        (NCCL: ncclAvg inside the collective; gloo has no AVG, so SUM then one scale).  Returns the averaged
        (xyz_loss, loss) view.  No-op without an initialised process group."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return self.metrics
        if dist.get_backend(group) == "nccl":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=group)
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.mul_(1.0 / dist.get_world_size(group))
        return self.metrics

    def overlap_plan(self):
        """Contiguous slices of the flat buffer in the order the backward finishes them (model/train_path.py:run): the upper
        half of the trunk blocks, the lower half, then everything else (decoder, head, shape encoder, tokens + the loss
        metrics in the tail).  global_transformer_blocks.* and local_transformer_blocks.* are each contiguous in
        named_parameters() order; block i of either list is final once the backward has passed global block i."""
        g0 = self.offsets["global_transformer_blocks.0.norm1.weight"]
        l0 = self.offsets["local_transformer_blocks.0.norm1.weight"]
        end = self.offsets["transformer_input_layernorm.weight"]
        nb = sum(1 for n in self.names if n.startswith("global_transformer_blocks.") and n.endswith(".norm1.weight"))
        per = (l0 - g0) // nb
        assert g0 + nb * per == l0 and l0 + nb * per == end, "trunk block slices are not contiguous"
        half = nb // 2
        return {"trunk_hi": [(g0 + half * per, l0), (l0 + half * per, end)],
                "trunk_lo": [(g0, g0 + half * per), (l0, l0 + half * per)],
                "rest": [(0, g0), (end, self.flat.numel())], "half": half}

    def packed_kv(self, prefix, d):
        """to_k.weight and to_v.weight gradients as one [2d, d] matrix (the forward runs them as one GEMM)."""
        ok, ov = self.offsets[prefix + "attn.to_k.weight"], self.offsets[prefix + "attn.to_v.weight"]
        assert ov == ok + d * d, "to_k / to_v gradient slices must be adjacent"
        return self.flat[ok:ok + 2 * d * d].view(2 * d, d)


def _wgrad(dY, X, n_out, k_in, rows, out32, alpha, ldy=None, ldx=None, ldo=None):
    """out32[n_out, k_in] += alpha * dY[rows, n_out]^T . X[rows, k_in]  (fp16 operands, row-major, split over rows)."""
    ops.gemm(dY, X, n_out, k_in, rows, lda=ldy if ldy is not None else dY.stride(0), ldw=ldx if ldx is not None else X.stride(0),
             tn=1, ksplit=_ksplit(n_out, k_in, rows), accumulate=1, out32=out32, ldo32=ldo if ldo is not None else k_in,
             out_scale=alpha)


class OverlappedAllReduce:
    """The gradient exchange of train.py's DDP (C1) overlapped with the hand-written backward: as soon as a contiguous part of
    the flat buffer is final, its all-reduce (AVG) is enqueued on a communication stream behind an event of the compute stream;
    ``finish()`` makes the compute stream wait for all of them.  Three waves (GradBuffer.overlap_plan): only the last one --
    the shape encoder's 176 MB -- has no backward left to hide behind."""

    def __init__(self, gb, group=None):
        import torch.distributed as dist
        self.gb, self.group, self.plan, self.works = gb, group, gb.overlap_plan(), []
        self.comm = torch.cuda.Stream(device=gb.flat.device) if gb.flat.is_cuda else None    # host buffers (gloo tests): no streams
        self.avg = dist.get_backend(group) == "nccl"
        self.world = dist.get_world_size(group)

    def _issue(self, tag):
        import torch.distributed as dist
        for lo, hi in self.plan[tag]:
            if hi > lo:
                self.works.append(dist.all_reduce(self.gb.flat[lo:hi], op=dist.ReduceOp.AVG if self.avg else dist.ReduceOp.SUM,
                                                  group=self.group, async_op=True))

    def ready(self, tag):
        if self.comm is None:
            self._issue(tag)
            return
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(self.comm):
            self.comm.wait_event(ev)
            self._issue(tag)

    def finish(self):
        for w in self.works:
            w.wait()
        self.works = []
        if not self.avg:
            self.gb.flat.mul_(1.0 / self.world)
        return self.gb.metrics


class TrainPath:
    def __init__(self, model):
        self.m = model
        self.grads = None
        self.step_id = 0
        self.on_ready = None     # callable(tag) or None: "trunk_hi" / "trunk_lo" / "rest" (GradBuffer.overlap_plan)

    # ------------------------------------------------------------------ buffers
    def buf(self, name, shape, dtype):
        return self.m._buf("tr." + name, shape, dtype)

    def grad_buffer(self):
        if self.grads is None or self.grads.flat.device != self.m.pos_embed.device:
            self.grads = GradBuffer(self.m)
        return self.grads

    # ------------------------------------------------------------------ self-attention block (transformer.py:379-423)
    def self_block_fwd(self, x_in, rows, Batt, L, w, tag):
        m, d, H = self.m, self.m.d, self.m.H
        sv = dict(x_in=x_in, rows=rows, Batt=Batt, L=L)
        sv["h1"] = h1 = self.buf(tag + ".h1", (rows, d), F16)
        sv["qkv"] = qkv = self.buf(tag + ".qkv", (rows, 3 * d), F16)
        sv["rstd"] = rstd = self.buf(tag + ".rstd", (rows, 2 * H), F32)
        sv["o"] = o = self.buf(tag + ".o", (rows, d), F16)
        sv["lse"] = lse = self.buf(tag + ".lse", (rows, H), F32)
        sv["x_mid"] = x_mid = self.buf(tag + ".x_mid", (rows, d), F32)
        sv["h2"] = h2 = self.buf(tag + ".h2", (rows, d), F16)
        sv["pre"] = pre = self.buf(tag + ".pre", (rows, 4 * d), F16)
        sv["hid"] = hid = self.buf(tag + ".hid", (rows, 4 * d), F16)
        x_out = self.buf(tag + ".x_out", (rows, d), F32)
        ops.layernorm(x_in, w["n1"], None, 1e-5, rows, d, out16=h1, ldo16=d)
        ops.gemm(h1, w["qkv"], rows, 3 * d, d, out16=qkv, ldo16=3 * d, qn_w=w["qn"], kn_w=w["kn"], qk_eps=1e-5, qk_cols=d,
                 qk_rstd=rstd, ld_rstd=2 * H)
        ops.attention(qkv, qkv[:, d:], qkv[:, 2 * d:], o, B=Batt, H=H, Lq=L, Lk=L, q_ld=3 * d, k_ld=3 * d, v_ld=3 * d, o_ld=d,
                      q_rows=rows, kv_rows=rows, q_batch_rows=L, kv_batch_rows=L, scale=m.dh ** -0.5, lse=lse, lse_ld=H)
        ops.gemm(o, w["fc"], rows, d, d, resid=x_in, ldr=d, out32=x_mid, ldo32=d)
        ops.layernorm(x_mid, w["n2"], None, 1e-5, rows, d, out16=h2, ldo16=d)
        ops.gemm(h2, w["w1"], rows, 4 * d, d, act=1, out16=hid, ldo16=4 * d, aux16=pre, ldaux=4 * d, aux_mode=1)
        ops.gemm(hid, w["w2"], rows, d, 4 * d, resid=x_mid, ldr=d, out32=x_out, ldo32=d)
        return x_out, sv

    def mlp_bwd(self, rows, dx32, dx16, x_mid, h2, pre, hid, w, wt, g, alpha, scratch):
        """x_out = x_mid + W2 gelu(W1 LN(x_mid)): dx (in place) <- dx + LN'(...); weight gradients accumulated."""
        d = self.m.d
        dU16, dy32 = scratch["dU16"], scratch["dy32"]
        ops.gemm(dx16, wt["w2"], rows, 4 * d, d, out16=dU16, ldo16=4 * d, aux16=pre, ldaux=4 * d, aux_mode=2)
        _wgrad(dx16, hid, d, 4 * d, rows, g["w2"], alpha)
        ops.gemm(dU16, wt["w1"], rows, d, 4 * d, out32=dy32, ldo32=d)
        _wgrad(dU16, h2, 4 * d, d, rows, g["w1"], alpha)
        ops.layernorm_bwd(dy32, x_mid, w["n2"], 1e-5, rows, d, dres=dx32, lddres=d, dx32=dx32, lddx32=d, dx16=dx16, lddx16=d,
                          dgamma=g["n2"], alpha=alpha)

    def self_block_bwd(self, sv, w, wt, g, dx32, dx16, alpha, scratch):
        m, d, H = self.m, self.m.d, self.m.H
        rows, Batt, L = sv["rows"], sv["Batt"], sv["L"]
        self.mlp_bwd(rows, dx32, dx16, sv["x_mid"], sv["h2"], sv["pre"], sv["hid"], w, wt, g, alpha, scratch)
        dO16, dy32, D, dqkv32, dqkv16 = scratch["dO16"], scratch["dy32"], scratch["D"], scratch["dqkv32"], scratch["dqkv16"]
        qkv = sv["qkv"]
        ops.gemm(dx16, wt["fc"], rows, d, d, out16=dO16, ldo16=d)
        _wgrad(dx16, sv["o"], d, d, rows, g["fc"], alpha)
        ops.attn_dot(dO16, d, sv["o"], d, rows, H, D, H)
        dqkv32[:rows, :d].zero_()                       # dQ accumulates over the K/V tiles (TMA reduce-add)
        ops.attention_bwd(qkv, qkv[:, d:], qkv[:, 2 * d:], dO16, sv["lse"], D, dqkv32, dqkv32[:, d:], dqkv32[:, 2 * d:], B=Batt, H=H,
                          Lq=L, Lk=L, q_ld=3 * d, k_ld=3 * d, v_ld=3 * d, do_ld=d, lse_ld=H, d_ld=H, dq_ld=3 * d, dk_ld=3 * d,
                          dv_ld=3 * d, q_rows=rows, kv_rows=rows, q_batch_rows=L, kv_batch_rows=L, scale=m.dh ** -0.5)
        ops.qknorm_bwd(dqkv32, 3 * d, qkv, 3 * d, sv["rstd"], 2 * H, w["qn"], w["kn"], d, 2 * d, 3 * d, rows, dqkv16, 3 * d,
                       g["qn"], g["kn"], alpha)
        ops.gemm(dqkv16, wt["qkv"], rows, d, 3 * d, out32=dy32, ldo32=d)
        _wgrad(dqkv16, sv["h1"], 3 * d, d, rows, g["qkv"], alpha)
        ops.layernorm_bwd(dy32, sv["x_in"], w["n1"], 1e-5, rows, d, dres=dx32, lddres=d, dx32=dx32, lddx32=d, dx16=dx16, lddx16=d,
                          dgamma=g["n1"], alpha=alpha)

    def scratch(self, rows, rows_qkv):
        """Gradient scratch shared by all blocks: `rows` covers the decoder chunk, `rows_qkv` the self-attention blocks."""
        d, H = self.m.d, self.m.H
        return dict(dU16=self.buf("s.dU16", (rows, 4 * d), F16), dy32=self.buf("s.dy32", (rows, d), F32),
                    dO16=self.buf("s.dO16", (rows, d), F16), D=self.buf("s.D", (rows, H), F32),
                    dqkv32=self.buf("s.dqkv32", (rows_qkv, 3 * d), F32), dqkv16=self.buf("s.dqkv16", (rows_qkv, 3 * d), F16))

    # ------------------------------------------------------------------ point features (Pcd_motion.py:456-459, 550-553)
    def point_features_fwd(self, P, xyz, normal, rgb, n, tag):
        d = self.m.d
        a0 = self.buf(tag + ".a0", (n, 2 * KP_EMB), F16)
        a1 = self.buf(tag + ".a1", (n, 2 * KP_FEAT), F16)
        feat = self.buf(tag + ".feat", (n, d), F32)
        ops.point_embed_features(xyz, n, a0, 2 * KP_EMB, KP_EMB)
        ops.gemm(a0, P["pe_w"], n, d, KP_EMB, passes=3, a_lo_off=KP_EMB, w_lo_off=KP_EMB, bias=P["pe_b"], out16=a1,
                 ldo16=2 * KP_FEAT, out16_lo_off=KP_FEAT)
        ops.point_extra_features(normal, rgb, n, a1, 2 * KP_FEAT, d, KP_FEAT, KP_FEAT)
        ops.gemm(a1, P["pn_w"], n, d, KP_FEAT, passes=3, a_lo_off=KP_FEAT, w_lo_off=KP_FEAT, bias=P["pn_b"], out32=feat, ldo32=d)
        return feat, a0, a1

    def point_features_bwd(self, PT, G, dfeat16, n, a0, a1, alpha):
        """feat = W_pn [emb | normal | rgb] + b_pn, emb = W_pe [sin | cos | xyz] + b_pe: parameter gradients only (inputs are data)."""
        d = self.m.d
        tmp_pn = self.buf("pf.tmp_pn", (d, KP_FEAT), F32)
        tmp_pe = self.buf("pf.tmp_pe", (d, KP_EMB), F32)
        da1 = self.buf("pf.da1", (n, d), F16)
        ops.colsum(dfeat16, d, n, d, G["point_normal_rgb_proj.bias"], alpha)
        tmp_pn.zero_()
        _wgrad(dfeat16, a1, d, KP_FEAT, n, tmp_pn, alpha, ldx=2 * KP_FEAT)
        ops.add_block(tmp_pn, KP_FEAT, d, d + 6, 1.0, 1, G["point_normal_rgb_proj.weight"], d + 6)
        ops.gemm(dfeat16, PT["pn"], n, d, d, out16=da1, ldo16=d)      # d emb = d feat . W_pn[:, :768]
        ops.colsum(da1, d, n, d, G["point_embed.mlp.bias"], alpha)
        tmp_pe.zero_()
        _wgrad(da1, a0, d, KP_EMB, n, tmp_pe, alpha, ldx=2 * KP_EMB)
        ops.add_block(tmp_pe, KP_EMB, d, 51, 1.0, 1, G["point_embed.mlp.weight"], 51)

    # ------------------------------------------------------------------ the step
    def run(self, sample, zero_grads=True, grad_scale=1.0):
        """One forward + backward.  Returns (pcd_moved [B,T,N,3], loss[2] = (mse, weight * mse)); parameter gradients of
        grad_scale * loss are accumulated into the flat gradient buffer."""
        m = self.m
        d, H, dh, M = m.d, m.H, m.dh, m.num_learnable_tokens
        scale = dh ** -0.5
        P = m._pack()
        PT = m._pack_transposed()
        GB = self.grad_buffer()
        G = GB.views
        if zero_grads:
            GB.flat.zero_()
        self.step_id += 1
        f32c = lambda t: t.detach().float().contiguous()
        ref_pcd = sample["ref_pcd"]
        dev = ref_pcd.device
        B, N = ref_pcd.shape[:2]
        S = sample["ref_shape_pcd"].shape[1]
        rgb_video = sample["rgb_video"]
        T, Hin, Win = rgb_video.shape[1:4]
        Fr = B * T
        if "point_clouds" not in sample:
            raise ValueError("training needs the ground-truth 'point_clouds' [B,T,N,3] in the sample (Pcd_motion.py:582-592)")
        target = f32c(sample["point_clouds"])
        if tuple(target.shape) != (B, T, N, 3):  # model/loss.py:50-57
            raise ValueError("Shape mismatch or invalid shape for coordinate MSE. Expected both tensors of shape (B, T, N, C). "
                             f"Got pred: {(B, T, N, 3)}, target: {tuple(target.shape)}")
        weight = float(m.config.training.coord_mse_loss_weight)
        alpha = grad_scale * 2.0 * weight / (float(B) * T * N * 3)

        def block_grads(pfx):
            return dict(n1=G[pfx + "norm1.weight"], qkv=G[pfx + "attn.to_qkv.weight"], fc=G[pfx + "attn.fc.weight"],
                        qn=G[pfx + "attn.q_norm.weight"], kn=G[pfx + "attn.k_norm.weight"], n2=G[pfx + "norm2.weight"],
                        w1=G[pfx + "mlp.mlp.0.weight"], w2=G[pfx + "mlp.mlp.2.weight"])

        def cross_grads(pfx):
            return dict(nq=G[pfx + "norm_q.weight"], nkv=G[pfx + "norm_kv.weight"], q=G[pfx + "attn.to_q.weight"],
                        kv=GB.packed_kv(pfx, d), fc=G[pfx + "attn.fc.weight"], qn=G[pfx + "attn.q_norm.weight"],
                        kn=G[pfx + "attn.k_norm.weight"], n2=G[pfx + "norm2.weight"], w1=G[pfx + "mlp.mlp.0.weight"],
                        w2=G[pfx + "mlp.mlp.2.weight"])

        # ================================================================== forward
        # ---- A. shape encoder (Pcd_motion.py:456-464)
        nS = B * S
        shape_feat, s_a0, s_a1 = self.point_features_fwd(P, f32c(sample["ref_shape_pcd"]).reshape(-1, 3),
                                                        f32c(sample["ref_shape_normals"]).reshape(-1, 3),
                                                        f32c(sample["ref_shape_rgbs"]).reshape(-1, 3), nS, "enc")
        e = P["enc"]
        e_qn16 = self.buf("enc.qn16", (M, d), F16)
        e_q16 = self.buf("enc.q16", (M, d), F16)
        e_qrstd = self.buf("enc.qrstd", (M, 2 * H), F32)
        e_kn16 = self.buf("enc.kn16", (nS, d), F16)
        e_kv16 = self.buf("enc.kv16", (nS, 2 * d), F16)
        e_kvrstd = self.buf("enc.kvrstd", (nS, 2 * H), F32)
        e_o16 = self.buf("enc.o16", (B * M, d), F16)
        e_lse = self.buf("enc.lse", (B * M, H), F32)
        mesh0 = self.buf("enc.mesh0", (B * M, d), F32)
        e_h2 = self.buf("enc.h2", (B * M, d), F16)
        e_pre = self.buf("enc.pre", (B * M, 4 * d), F16)
        e_hid = self.buf("enc.hid", (B * M, 4 * d), F16)
        mesh = self.buf("enc.mesh1", (B * M, d), F32)
        ops.layernorm(P["tok"], e["nq"], None, 1e-5, M, d, out16=e_qn16, ldo16=d)
        ops.gemm(e_qn16, e["q"], M, d, d, out16=e_q16, ldo16=d, qn_w=e["qn"], kn_w=None, qk_cols=d, qk_rstd=e_qrstd, ld_rstd=2 * H)
        ops.layernorm(shape_feat, e["nkv"], None, 1e-5, nS, d, out16=e_kn16, ldo16=d)
        ops.gemm(e_kn16, e["kv"], nS, 2 * d, d, out16=e_kv16, ldo16=2 * d, qn_w=e["kn"], kn_w=None, qk_cols=d, qk_rstd=e_kvrstd,
                 ld_rstd=2 * H)
        ops.attention(e_q16, e_kv16, e_kv16[:, d:], e_o16, B=B, H=H, Lq=M, Lk=S, q_ld=d, k_ld=2 * d, v_ld=2 * d, o_ld=d, q_rows=M,
                      kv_rows=nS, q_batch_rows=0, kv_batch_rows=S, scale=scale, lse=e_lse, lse_ld=H)
        ops.gemm(e_o16, e["fc"], B * M, d, d, resid=P["tok"], ldr=d, resid_mod=M, out32=mesh0, ldo32=d)
        ops.layernorm(mesh0, e["n2"], None, 1e-5, B * M, d, out16=e_h2, ldo16=d)
        ops.gemm(e_h2, e["w1"], B * M, 4 * d, d, act=1, out16=e_hid, ldo16=4 * d, aux16=e_pre, ldaux=4 * d, aux_mode=1)
        ops.gemm(e_hid, e["w2"], B * M, d, 4 * d, resid=mesh0, ldr=d, out32=mesh, ldo32=d)
        pts_saved = []
        for i, w in enumerate(P["pts"]):
            mesh, sv = self.self_block_fwd(mesh, B * M, B, M, w, f"pts{i}")
            pts_saved.append(sv)

        # ---- B. frozen DINOv2 (no gradient, nothing kept)
        xd, npatch = m._dino_forward(P, rgb_video, Fr, Hin, Win)

        # ---- C. token assembly (Pcd_motion.py:477-509); pos_drop is active in train() (:369-370, 490)
        L = 4 + M + npatch
        rows_t = Fr * L
        x0 = self.buf("trunk.x0", (rows_t, d), F32)
        tok_pre = self.buf("trunk.pre_ln", (rows_t, d), F32)
        drop_p = float(m.drop_rate) if m.training else 0.0
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if drop_p > 0 else 0
        ops.assemble_tokens(xd, P["d_nw"], P["d_nb"], 1e-6, m._pos_for(T), P["sp0"], P["spr"], mesh, P["in_ln"], 1e-5, B, T, M,
                            npatch, d, x0, drop_p=drop_p, seed=seed, pre_out=tok_pre)

        # ---- D. alternating global / local attention (Pcd_motion.py:394-409)
        x = x0
        trunk_saved = []
        for i, (wg, wl) in enumerate(zip(P["glb"], P["loc"])):
            x, sv = self.self_block_fwd(x, rows_t, B, T * L, wg, f"glb{i}")
            trunk_saved.append(("glb", i, sv))
            x, sv = self.self_block_fwd(x, rows_t, Fr, L, wl, f"loc{i}")
            trunk_saved.append(("loc", i, sv))

        # ---- E. decoder (Pcd_motion.py:520-579): forward + backward per frame chunk
        dc, dcT, gd = P["dec"], PT["dec"], cross_grads("decoder_cross_attn.")
        nN = B * N
        feat, d_a0, d_a1 = self.point_features_fwd(P, f32c(ref_pcd).reshape(-1, 3), f32c(sample["ref_normal"]).reshape(-1, 3),
                                                   f32c(sample["ref_rgb"]).reshape(-1, 3), nN, "dec")
        dqn16 = self.buf("dec.qn16", (nN, d), F16)
        dq16 = self.buf("dec.q16", (nN, d), F16)
        dq_rstd = self.buf("dec.qrstd", (nN, 2 * H), F32)
        ops.layernorm(feat, dc["nq"], None, 1e-5, nN, d, out16=dqn16, ldo16=d)
        ops.gemm(dqn16, dc["q"], nN, d, d, out16=dq16, ldo16=d, qn_w=dc["qn"], kn_w=None, qk_cols=d, qk_rstd=dq_rstd, ld_rstd=2 * H)
        dkn16 = self.buf("dec.kn16", (Fr * M, d), F16)
        dkv16 = self.buf("dec.kv16", (Fr * M, 2 * d), F16)
        dkv_rstd = self.buf("dec.kvrstd", (Fr * M, 2 * H), F32)
        ops.layernorm(x, dc["nkv"], None, 1e-5, Fr * M, d, src_rpg=M, src_gstride=L, src_goff=4, out16=dkn16, ldo16=d)
        ops.gemm(dkn16, dc["kv"], Fr * M, 2 * d, d, out16=dkv16, ldo16=2 * d, qn_w=dc["kn"], kn_w=None, qk_cols=d, qk_rstd=dkv_rstd,
                 ld_rstd=2 * H)

        out = torch.empty(B, T, N, 3, device=dev, dtype=F32)
        partials = self.buf("mse_partials", (max(Fr, 1) * 1024,), F32)
        dQ32 = self.buf("dec.dQ32", (nN, d), F32)
        dfeat32 = self.buf("dec.dfeat32", (nN, d), F32)
        dKV32 = self.buf("dec.dKV32", (Fr * M, 2 * d), F32)
        dQ32.zero_()
        dfeat32.zero_()
        n_part = 0
        tchunk = max(1, min(T, m.max_decode_rows // max(N, 1)))
        R = tchunk * N
        o16 = self.buf("dec.o16", (R, d), F16)
        lse = self.buf("dec.lse", (R, H), F32)
        xdec = self.buf("dec.x", (R, d), F32)
        h2 = self.buf("dec.h2", (R, d), F16)
        pre = self.buf("dec.pre", (R, 4 * d), F16)
        hid = self.buf("dec.hid", (R, 4 * d), F16)
        xdec2 = self.buf("dec.x2", (R, d), F32)
        hl = self.buf("dec.hl", (R, 2 * d), F16)
        u32 = self.buf("dec.u32", (R, d), F32)
        du16 = self.buf("dec.du16", (R, d), F16)
        dx32 = self.buf("dec.dx32", (R, d), F32)
        dx16 = self.buf("dec.dx16", (R, d), F16)
        sc = self.scratch(max(R, rows_t, B * M), max(rows_t, B * M))
        for b in range(B):
            for t0 in range(0, T, tchunk):
                tc = min(tchunk, T - t0)
                rows = tc * N
                f0 = b * T + t0
                qb, kvb = dq16[b * N:], dkv16[f0 * M:]
                ops.attention(qb, kvb, kvb[:, d:], o16, B=tc, H=H, Lq=N, Lk=M, q_ld=d, k_ld=2 * d, v_ld=2 * d, o_ld=d, q_rows=N,
                              kv_rows=tc * M, q_batch_rows=0, kv_batch_rows=M, scale=scale, lse=lse, lse_ld=H)
                ops.gemm(o16, dc["fc"], rows, d, d, resid=feat[b * N:], ldr=d, resid_mod=N, out32=xdec, ldo32=d)
                ops.layernorm(xdec, dc["n2"], None, 1e-5, rows, d, out16=h2, ldo16=d)
                ops.gemm(h2, dc["w1"], rows, 4 * d, d, act=1, out16=hid, ldo16=4 * d, aux16=pre, ldaux=4 * d, aux_mode=1)
                ops.gemm(hid, dc["w2"], rows, d, 4 * d, resid=xdec, ldr=d, out32=xdec2, ldo32=d)
                # shared_mlp_output (Pcd_motion.py:336-341, 561); the pre-activation of its GELU is kept in fp32
                ops.layernorm(xdec2, P["h_lnw"], P["h_lnb"], 1e-5, rows, d, out16=hl, ldo16=2 * d, lo_off=d)
                ops.gemm(hl, P["h1_w"], rows, d, d, passes=3, a_lo_off=d, w_lo_off=d, bias=P["h1_b"], out32=u32, ldo32=d)
                o_view, tgt = out[b, t0:t0 + tc], target[b, t0:t0 + tc]
                n_part += ops.head3_mse(u32, d, P["h3_w"], P["h3_b"], rows, d, o_view, tgt, partials[n_part:], pre_gelu=1)
                # ---------------- backward of this chunk (autograd of Pcd_motion.py:556-562 + model/loss.py:59-61)
                ops.head_bwd(o_view, tgt, u32, d, P["h3_w"], rows, d, du16, d, G["shared_mlp_output.3.weight"],
                             G["shared_mlp_output.3.bias"], alpha)
                ops.colsum(du16, d, rows, d, G["shared_mlp_output.1.bias"], alpha)
                _wgrad(du16, hl, d, d, rows, G["shared_mlp_output.1.weight"], alpha, ldx=2 * d)
                ops.gemm(du16, PT["h1"], rows, d, d, out32=sc["dy32"], ldo32=d)
                ops.layernorm_bwd(sc["dy32"], xdec2, P["h_lnw"], 1e-5, rows, d, dx32=dx32, lddx32=d, dx16=dx16, lddx16=d,
                                  dgamma=G["shared_mlp_output.0.weight"], dbeta=G["shared_mlp_output.0.bias"], alpha=alpha)
                self.mlp_bwd(rows, dx32, dx16, xdec, h2, pre, hid, dc, dcT, gd, alpha, sc)
                ops.gemm(dx16, dcT["fc"], rows, d, d, out16=sc["dO16"], ldo16=d)
                _wgrad(dx16, o16, d, d, rows, gd["fc"], alpha)
                # the residual of every frame is the same per-point feature (Pcd_motion.py:556-560): sum over the frames
                ops.sum_groups(dx32, d, tc, N, N, d, accumulate=1, out32=dfeat32[b * N:], ldo32=d)
                ops.attn_dot(sc["dO16"], d, o16, d, rows, H, sc["D"], H)
                dkvb = dKV32[f0 * M:]
                ops.attention_bwd(qb, kvb, kvb[:, d:], sc["dO16"], lse, sc["D"], dQ32[b * N:], dkvb, dkvb[:, d:], B=tc, H=H, Lq=N,
                                  Lk=M, q_ld=d, k_ld=2 * d, v_ld=2 * d, do_ld=d, lse_ld=H, d_ld=H, dq_ld=d, dk_ld=2 * d, dv_ld=2 * d,
                                  q_rows=N, kv_rows=tc * M, q_batch_rows=0, kv_batch_rows=M, scale=scale)
        loss = GB.metrics      # (mse, weight * mse) live in the tail of the flat gradient buffer: they ride in its all-reduce
        ops.mse_finalize(partials, n_part, float(B) * T * N * 3, weight, loss)

        # ================================================================== backward
        # ---- decoder query side: q-norm, to_q, norm_q, residual -> point features
        dqraw16 = self.buf("dec.dqraw16", (nN, d), F16)
        dfeat16 = self.buf("dec.dfeat16", (nN, d), F16)
        dyN = self.buf("dec.dyN", (nN, d), F32)
        ops.qknorm_bwd(dQ32, d, dq16, d, dq_rstd, 2 * H, dc["qn"], None, d, d, d, nN, dqraw16, d, gd["qn"], None, alpha)
        _wgrad(dqraw16, dqn16, d, d, nN, gd["q"], alpha)
        ops.gemm(dqraw16, dcT["q"], nN, d, d, out32=dyN, ldo32=d)
        ops.layernorm_bwd(dyN, feat, dc["nq"], 1e-5, nN, d, dres=dfeat32, lddres=d, dx16=dfeat16, lddx16=d, dgamma=gd["nq"],
                          alpha=alpha)
        self.point_features_bwd(PT, G, dfeat16, nN, d_a0, d_a1, alpha)
        # ---- decoder key / value side -> trunk output rows 4..4+M of every frame (Pcd_motion.py:520)
        nKV = Fr * M
        dkvraw16 = self.buf("dec.dkvraw16", (nKV, 2 * d), F16)
        dyKV = self.buf("dec.dyKV", (nKV, d), F32)
        ops.qknorm_bwd(dKV32, 2 * d, dkv16, 2 * d, dkv_rstd, 2 * H, dc["kn"], None, d, d, 2 * d, nKV, dkvraw16, 2 * d, gd["kn"], None,
                       alpha)
        _wgrad(dkvraw16, dkn16, 2 * d, d, nKV, gd["kv"], alpha)
        ops.gemm(dkvraw16, dcT["kv"], nKV, d, 2 * d, out32=dyKV, ldo32=d)
        tx32 = self.buf("trunk.dx32", (rows_t, d), F32)
        tx16 = self.buf("trunk.dx16", (rows_t, d), F16)
        tx32.zero_()
        tx16.zero_()
        ops.layernorm_bwd(dyKV, x, dc["nkv"], 1e-5, nKV, d, src_rpg=M, src_gstride=L, src_goff=4, dx32=tx32, lddx32=d, dx16=tx16,
                          lddx16=d, dgamma=gd["nkv"], alpha=alpha)
        # ---- trunk
        half = GB.overlap_plan()["half"] if self.on_ready is not None else -1
        for kind, i, sv in reversed(trunk_saved):
            name = "global_transformer_blocks" if kind == "glb" else "local_transformer_blocks"
            self.self_block_bwd(sv, P[kind][i], PT[kind][i], block_grads(f"{name}.{i}."), tx32, tx16, alpha, sc)
            if self.on_ready is not None and kind == "glb" and i in (half, 0):    # global block i is the last of pair i in backward order
                self.on_ready("trunk_hi" if i == half and half > 0 else "trunk_lo")
        # ---- token assembly: transformer_input_layernorm over every token; gradients of the special tokens and of the mesh
        # tokens (broadcast to all frames, Pcd_motion.py:495-507) are sums over frames.  DINOv2 is frozen: its rows stop here.
        dtok = self.buf("trunk.dtok", (rows_t, d), F32)
        ops.layernorm_bwd(tx32, tok_pre, P["in_ln"], 1e-5, rows_t, d, dx32=dtok, lddx32=d,
                          dgamma=G["transformer_input_layernorm.weight"], alpha=alpha)
        ops.sum_groups(dtok, d, B, T * L, 4, d, scale=alpha, accumulate=1, out32=G["special_token_0"].view(4, d), ldo32=d)
        ops.sum_groups(dtok, d, Fr, L, 4, d, scale=alpha, accumulate=1, out32=G["special_token_rest"].view(4, d), ldo32=d)
        ops.sum_groups(dtok, d, B, T * L, 4, d, scale=-alpha, accumulate=1, out32=G["special_token_rest"].view(4, d), ldo32=d)
        mx32 = self.buf("enc.dx32", (B * M, d), F32)
        mx16 = self.buf("enc.dx16", (B * M, d), F16)
        ops.sum_groups(dtok, d, T, L, B * M, d, rpg=M, in_gstride=T * L, in_goff=4, out32=mx32, ldo32=d, out16=mx16, ldo16=d)
        # ---- shape encoder: 4 self-attention blocks, then the cross-attention block into the learnable tokens
        for i in reversed(range(len(pts_saved))):
            self.self_block_bwd(pts_saved[i], P["pts"][i], PT["pts"][i], block_grads(f"points_transformer_blocks.{i}."), mx32, mx16,
                                alpha, sc)
        eT, ge = PT["enc"], cross_grads("encoder_cross_attn.")
        self.mlp_bwd(B * M, mx32, mx16, mesh0, e_h2, e_pre, e_hid, e, eT, ge, alpha, sc)
        ops.gemm(mx16, eT["fc"], B * M, d, d, out16=sc["dO16"], ldo16=d)
        _wgrad(mx16, e_o16, d, d, B * M, ge["fc"], alpha)
        dtokq = self.buf("enc.dtokq", (M, d), F32)          # gradient of the learnable tokens (residual + query path)
        ops.sum_groups(mx32, d, B, M, M, d, out32=dtokq, ldo32=d)
        ops.attn_dot(sc["dO16"], d, e_o16, d, B * M, H, sc["D"], H)
        e_dQ = self.buf("enc.dQ32", (M, d), F32)
        e_dKV = self.buf("enc.dKV32", (nS, 2 * d), F32)
        e_dQ.zero_()
        ops.attention_bwd(e_q16, e_kv16, e_kv16[:, d:], sc["dO16"], e_lse, sc["D"], e_dQ, e_dKV, e_dKV[:, d:], B=B, H=H, Lq=M, Lk=S,
                          q_ld=d, k_ld=2 * d, v_ld=2 * d, do_ld=d, lse_ld=H, d_ld=H, dq_ld=d, dk_ld=2 * d, dv_ld=2 * d, q_rows=M,
                          kv_rows=nS, q_batch_rows=0, kv_batch_rows=S, scale=scale)
        e_dqraw = self.buf("enc.dqraw16", (M, d), F16)
        e_dyq = self.buf("enc.dyq", (M, d), F32)
        ops.qknorm_bwd(e_dQ, d, e_q16, d, e_qrstd, 2 * H, e["qn"], None, d, d, d, M, e_dqraw, d, ge["qn"], None, alpha)
        _wgrad(e_dqraw, e_qn16, d, d, M, ge["q"], alpha)
        ops.gemm(e_dqraw, eT["q"], M, d, d, out32=e_dyq, ldo32=d)
        ops.layernorm_bwd(e_dyq, P["tok"], e["nq"], 1e-5, M, d, dres=dtokq, lddres=d, dx32=dtokq, lddx32=d, dgamma=ge["nq"], alpha=alpha)
        ops.add_block(dtokq, d, M, d, alpha, 1, G["learnable_tokens"].view(M, d), d)
        e_dkvraw = self.buf("enc.dkvraw16", (nS, 2 * d), F16)
        e_dykv = self.buf("enc.dykv", (nS, d), F32)
        e_dfeat16 = self.buf("enc.dfeat16", (nS, d), F16)
        ops.qknorm_bwd(e_dKV, 2 * d, e_kv16, 2 * d, e_kvrstd, 2 * H, e["kn"], None, d, d, 2 * d, nS, e_dkvraw, 2 * d, ge["kn"], None,
                       alpha)
        _wgrad(e_dkvraw, e_kn16, 2 * d, d, nS, ge["kv"], alpha)
        ops.gemm(e_dkvraw, eT["kv"], nS, d, 2 * d, out32=e_dykv, ldo32=d)
        ops.layernorm_bwd(e_dykv, shape_feat, e["nkv"], 1e-5, nS, d, dx16=e_dfeat16, lddx16=d, dgamma=ge["nkv"], alpha=alpha)
        self.point_features_bwd(PT, G, e_dfeat16, nS, s_a0, s_a1, alpha)
        if self.on_ready is not None:
            self.on_ready("rest")
        return out, loss
