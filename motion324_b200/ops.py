"""Thin torch-tensor wrappers over the C ABI (pointer + size extraction only; all compute is in libm324.so)."""
import ctypes as C

import torch

from . import lib as _l


LAUNCHES = [0]   # kernels launched through this module (bench.py reads it)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _chk_f32(*ts):
    for t in ts:
        if t is not None:
            assert t.is_cuda and t.dtype == torch.float32, (t.device, t.dtype)


def _chk_f16(*ts):
    for t in ts:
        if t is not None:
            assert t.is_cuda and t.dtype == torch.float16, (t.device, t.dtype)


def launch_count():
    """Kernels launched by libm324 in this process (counted inside the library, one per cudaLaunchKernelEx)."""
    return int(_l.load().m324_launch_count())


def set_tuning(knob, value):
    _l.check(_l.load().m324_set_tuning(int(knob), int(value)), "m324_set_tuning")


def check_device():
    _l.check(_l.load().m324_check_device(), "m324_check_device")


def gemm(A, W, M, N, K, *, lda=None, ldw=None, passes=1, a_lo_off=0, w_lo_off=0, bias=None, gamma=None, resid=None,
         ldr=0, resid_mod=0, resid_div=0, out32=None, ldo32=0, out16=None, ldo16=0, out16_lo_off=0, act=0, qn_w=None,
         kn_w=None, qk_eps=1e-5, qk_cols=0, force_bn128=0, tn=0, ksplit=0, accumulate=0, aux16=None, ldaux=0, aux_mode=0,
         qk_rstd=None, ld_rstd=0, out_scale=0.0, head_w=None, head_part=None):
    """A, W: fp16 or bf16 tensors (or (tensor, element_offset) views resolved by the caller); see include/m324.h."""
    for t in (A, W, out16):
        assert t is None or (t.is_cuda and t.dtype in (torch.float16, torch.bfloat16)), (t.device, t.dtype)
    _chk_f16(aux16)
    _chk_f32(bias, gamma, resid, out32, qn_w, kn_w, qk_rstd, head_w, head_part)
    a = _l.GemmArgs()
    a.A, a.lda, a.W, a.ldw = A.data_ptr(), (lda if lda is not None else A.stride(0)), W.data_ptr(), (ldw if ldw is not None else W.stride(0))
    assert A.dtype == W.dtype, "tcgen05 kind::f16 takes both operands in the same format"
    a.M, a.N, a.K, a.passes, a.a_lo_off, a.w_lo_off, a.bf16 = M, N, K, passes, a_lo_off, w_lo_off, int(A.dtype == torch.bfloat16)
    a.bias = bias.data_ptr() if bias is not None else None
    a.gamma = gamma.data_ptr() if gamma is not None else None
    a.resid = resid.data_ptr() if resid is not None else None
    a.ldr, a.resid_mod, a.resid_div = ldr, resid_mod, resid_div
    a.out32 = out32.data_ptr() if out32 is not None else None
    a.ldo32 = ldo32
    a.out16 = out16.data_ptr() if out16 is not None else None
    a.ldo16, a.out16_lo_off, a.act = ldo16, out16_lo_off, act
    a.qn_w = qn_w.data_ptr() if qn_w is not None else None
    a.kn_w = kn_w.data_ptr() if kn_w is not None else None
    a.qk_eps, a.qk_cols, a.force_bn128 = qk_eps, qk_cols, force_bn128
    a.tn, a.ksplit, a.accumulate = tn, ksplit, accumulate
    a.aux16 = aux16.data_ptr() if aux16 is not None else None
    a.ldaux, a.aux_mode = ldaux, aux_mode
    a.out16_bf16 = int(out16 is not None and out16.dtype == torch.bfloat16)
    a.qk_rstd = qk_rstd.data_ptr() if qk_rstd is not None else None
    a.ld_rstd = ld_rstd
    a.out_scale = out_scale
    a.head_w = head_w.data_ptr() if head_w is not None else None
    a.head_part = head_part.data_ptr() if head_part is not None else None
    LAUNCHES[0] += 1
    _l.check(_l.load().m324_gemm(C.byref(a), _stream()), "m324_gemm")


_ATTN_WS = {}


def attention_workspace(device):
    """The caller-owned scratch m324_attention may use to split the work items of a partly filled last wave (one per device
    and stream of use; ~10 MB)."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _ATTN_WS.get(key)
    if ws is None:
        ws = _ATTN_WS[key] = torch.empty(int(_l.load().m324_attention_workspace_bytes()), dtype=torch.uint8, device=device)
    return ws


def attention(q, k, v, out, *, B, H, Lq, Lk, q_ld, k_ld, v_ld, o_ld, q_rows, kv_rows, q_batch_rows, kv_batch_rows,
              q_batch_div=1, scale=0.125, lse=None, lse_ld=0, workspace="auto", partial=None):
    """q/k/v/out: fp16 tensors whose data_ptr() is (row 0, head 0) of the operand.  workspace: "auto" (module-owned scratch),
    a uint8 tensor, or None (never split work items).  partial = (parts, index): this call is one K/V range of a multi-call
    attention (workspace of attention_partial_bytes() required, finish with attention_merge())."""
    _chk_f16(q, k, v, out)
    _chk_f32(lse)
    if isinstance(workspace, str):
        workspace = attention_workspace(q.device)
    a = _l.AttnArgs()
    a.q, a.q_ld, a.q_rows = q.data_ptr(), q_ld, q_rows
    a.k, a.k_ld = k.data_ptr(), k_ld
    a.v, a.v_ld, a.kv_rows = v.data_ptr(), v_ld, kv_rows
    a.B, a.H, a.Lq, a.Lk = B, H, Lq, Lk
    a.q_batch_rows, a.kv_batch_rows, a.q_batch_div = q_batch_rows, kv_batch_rows, q_batch_div
    a.out, a.o_ld, a.scale = out.data_ptr(), o_ld, scale
    a.lse, a.lse_ld = (lse.data_ptr() if lse is not None else None), lse_ld
    a.workspace, a.workspace_bytes = (workspace.data_ptr(), workspace.numel()) if workspace is not None else (None, 0)
    a.partial_parts, a.partial_index = partial if partial is not None else (0, 0)
    LAUNCHES[0] += 1
    _l.check(_l.load().m324_attention(C.byref(a), _stream()), "m324_attention")


def attention_partial_bytes(B, H, Lq, parts):
    return int(_l.load().m324_attention_partial_bytes(B, H, Lq, parts))


def attention_merge(out, *, B, H, Lq, o_ld, parts, workspace, lse=None, lse_ld=0):
    """Merge the `parts` partial attention() calls (log-sum-exp) into out [B*Lq, o_ld] (and lse)."""
    _chk_f16(out)
    _chk_f32(lse)
    a = _l.AttnArgs()
    a.B, a.H, a.Lq = B, H, Lq
    a.out, a.o_ld = out.data_ptr(), o_ld
    a.lse, a.lse_ld = (lse.data_ptr() if lse is not None else None), lse_ld
    a.workspace, a.workspace_bytes, a.partial_parts = workspace.data_ptr(), workspace.numel(), parts
    LAUNCHES[0] += 1
    _l.check(_l.load().m324_attention_merge(C.byref(a), _stream()), "m324_attention_merge")


def layernorm(x, w, b, eps, rows, cols, *, ldx=None, src_rpg=0, src_gstride=0, src_goff=0, out16=None, ldo16=0, lo_off=0,
              out32=None, ldo32=0):
    _chk_f32(x, w, b, out32)
    _chk_f16(out16)
    LAUNCHES[0] += 1
    _l.check(_l.load().m324_layernorm(_p(x), ldx if ldx is not None else cols, _p(w), _p(b), eps, rows, cols, src_rpg,
                                      src_gstride, src_goff, _p(out16), ldo16, lo_off, _p(out32), ldo32, _stream()),
             "m324_layernorm")


def point_embed_features(xyz, n, out, ldo, lo_off):
    _chk_f32(xyz); _chk_f16(out)
    LAUNCHES[0] += 1
    _l.check(_l.load().m324_point_embed_features(_p(xyz), n, _p(out), ldo, lo_off, _stream()), "m324_point_embed_features")


def point_extra_features(normal, rgb, n, out, ldo, col0, kpad, lo_off):
    _chk_f32(normal, rgb); _chk_f16(out)
    LAUNCHES[0] += 1
    _l.check(_l.load().m324_point_extra_features(_p(normal), _p(rgb), n, _p(out), ldo, col0, kpad, lo_off, _stream()),
             "m324_point_extra_features")


def preprocess_frames(video, F, Hin, Win, S, patches, ldp, kpad):
    _chk_f32(video); _chk_f16(patches)
    LAUNCHES[0] += 1
    _l.check(_l.load().m324_preprocess_frames(_p(video), F, Hin, Win, S, _p(patches), ldp, kpad, _stream()),
             "m324_preprocess_frames")


def dino_assemble(patch, cls, pos, F, np_, C_, x):
    _chk_f32(patch, cls, pos, x)
    LAUNCHES[0] += 1
    _l.check(_l.load().m324_dino_assemble(_p(patch), _p(cls), _p(pos), F, np_, C_, _p(x), _stream()), "m324_dino_assemble")


def assemble_tokens(dino_x, dino_nw, dino_nb, dino_eps, pos_embed, sp0, sprest, mesh_feat, ln_w, ln_eps, B, T, ntok, npatch,
                    C_, out, drop_p=0.0, seed=0, pre_out=None):
    _chk_f32(dino_x, dino_nw, dino_nb, pos_embed, sp0, sprest, mesh_feat, ln_w, out, pre_out)
    LAUNCHES[0] += 1
    _l.check(_l.load().m324_assemble_tokens(_p(dino_x), _p(dino_nw), _p(dino_nb), dino_eps, _p(pos_embed), _p(sp0),
                                            _p(sprest), _p(mesh_feat), _p(ln_w), ln_eps, B, T, ntok, npatch, C_, _p(out),
                                            float(drop_p), int(seed) & 0xFFFFFFFFFFFFFFFF, _p(pre_out), _stream()), "m324_assemble_tokens")


def head3_mse(h, ldh, w3, b3, rows, C_, out, target, partials, pre_gelu=0):
    _chk_f32(h, w3, b3, out, target, partials)
    n = C.c_int32(0)
    LAUNCHES[0] += 1
    _l.check(_l.load().m324_head3_mse(_p(h), ldh, _p(w3), _p(b3), rows, C_, _p(out), _p(target), _p(partials), C.byref(n),
                                      int(pre_gelu), _stream()), "m324_head3_mse")
    return n.value


def mse_finalize(partials, n, count, weight, loss):
    _chk_f32(partials, loss)
    LAUNCHES[0] += 1
    _l.check(_l.load().m324_mse_finalize(_p(partials), n, float(count), float(weight), _p(loss), _stream()), "m324_mse_finalize")


def mse_loss(pred, target, n, weight, partials, loss):
    _chk_f32(pred, target, partials, loss)
    LAUNCHES[0] += 2
    _l.check(_l.load().m324_mse_loss(_p(pred), _p(target), n, float(weight), _p(partials), _p(loss), _stream()), "m324_mse_loss")


def cast_pad_f16(src, rows, cols, dst, ldo, kpad, lo_off=0, lds=None):
    _chk_f32(src); _chk_f16(dst)
    LAUNCHES[0] += 1
    _l.check(_l.load().m324_cast_pad_f16(_p(src), lds if lds is not None else cols, rows, cols, _p(dst), ldo, kpad, lo_off,
                                         _stream()), "m324_cast_pad_f16")


def smooth_trajectories(trajs, out, motion_threshold, sigma, do_threshold, do_gaussian):
    _chk_f32(trajs, out)
    B, T, N, _ = trajs.shape
    LAUNCHES[0] += 1
    _l.check(_l.load().m324_smooth_trajectories(_p(trajs), _p(out), B, T, N, float(motion_threshold), float(sigma), int(do_threshold),
                                                int(do_gaussian), _stream()), "m324_smooth_trajectories")


def chamfer_nn(points1, points2, dist1, idx1, dist2, idx2):
    """points1 [F, n1, 3], points2 [F, n2, 3] (both fp32 or both fp64, CUDA); dist1/idx1 [F, n2], dist2/idx2 [F, n1]."""
    assert points1.is_cuda and points2.is_cuda and points1.dtype == points2.dtype and points1.dtype in (torch.float32, torch.float64)
    assert points1.is_contiguous() and points2.is_contiguous() and points1.shape[-1] == 3 and points2.shape[-1] == 3
    assert dist1.dtype == torch.float64 and dist2.dtype == torch.float64
    assert idx1 is None or idx1.dtype == torch.int32
    assert idx2 is None or idx2.dtype == torch.int32
    F, n1, _ = points1.shape
    n2 = points2.shape[1]
    assert points2.shape[0] == F
    LAUNCHES[0] += 1
    _l.check(_l.load().m324_chamfer_nn(_p(points1), n1, _p(points2), n2, F, int(points1.dtype == torch.float64), _p(dist1), _p(idx1),
                                       _p(dist2), _p(idx2), _stream()), "m324_chamfer_nn")


def chamfer_reduce(dist1, dist2, threshold, out):
    """dist1 [F, n2], dist2 [F, n1] fp64 -> out [F, 4] = (chamfer, fscore, precision, recall)."""
    assert dist1.dtype == torch.float64 and dist2.dtype == torch.float64 and out.dtype == torch.float64
    F, n2 = dist1.shape
    n1 = dist2.shape[1]
    LAUNCHES[0] += 1
    _l.check(_l.load().m324_chamfer_reduce(_p(dist1), n2, _p(dist2), n1, F, float(threshold), _p(out), _stream()), "m324_chamfer_reduce")


# ------------------------------------------------------------------------------------------------ backward
def layernorm_bwd(dy, x, w, eps, rows, cols, *, lddy=None, ldx=None, src_rpg=0, src_gstride=0, src_goff=0, dres=None, lddres=0,
                  dx32=None, lddx32=0, dx16=None, lddx16=0, dgamma=None, dbeta=None, alpha=1.0):
    _chk_f32(dy, x, w, dres, dx32, dgamma, dbeta)
    _chk_f16(dx16)
    LAUNCHES[0] += 1
    _l.check(_l.load().m324_layernorm_bwd(_p(dy), lddy if lddy is not None else cols, _p(x), ldx if ldx is not None else cols, _p(w), eps,
                                          rows, cols, src_rpg, src_gstride, src_goff, _p(dres), lddres, _p(dx32), lddx32, _p(dx16), lddx16,
                                          _p(dgamma), _p(dbeta), float(alpha), _stream()), "m324_layernorm_bwd")


def qknorm_bwd(d_in, ld_in, y16, ldy, rstd, ld_rstd, wq, wk, q_cols, norm_cols, cols, rows, out16, ldo, dwq, dwk, alpha):
    _chk_f32(d_in, rstd, wq, wk, dwq, dwk)
    _chk_f16(y16, out16)
    LAUNCHES[0] += 1
    _l.check(_l.load().m324_qknorm_bwd(_p(d_in), ld_in, _p(y16), ldy, _p(rstd), ld_rstd, _p(wq), _p(wk), q_cols, norm_cols, cols, rows,
                                       _p(out16), ldo, _p(dwq), _p(dwk), float(alpha), _stream()), "m324_qknorm_bwd")


def head_bwd(pred, target, u, ldu, w3, rows, C_, du16, lddu, dw3, db3, alpha):
    _chk_f32(pred, target, u, w3, dw3, db3)
    _chk_f16(du16)
    LAUNCHES[0] += 1
    _l.check(_l.load().m324_head_bwd(_p(pred), _p(target), _p(u), ldu, _p(w3), rows, C_, _p(du16), lddu, _p(dw3), _p(db3), float(alpha),
                                     _stream()), "m324_head_bwd")


def colsum(dy16, ld, rows, cols, db, alpha):
    _chk_f16(dy16); _chk_f32(db)
    LAUNCHES[0] += 1
    _l.check(_l.load().m324_colsum(_p(dy16), ld, rows, cols, _p(db), float(alpha), _stream()), "m324_colsum")


def sum_groups(inp, ld_in, ngroups, group_stride, rows, cols, *, rpg=0, in_gstride=0, in_goff=0, scale=1.0, accumulate=0, out32=None,
               ldo32=0, out16=None, ldo16=0):
    _chk_f32(inp, out32); _chk_f16(out16)
    LAUNCHES[0] += 1
    _l.check(_l.load().m324_sum_groups(_p(inp), ld_in, ngroups, group_stride, rpg, in_gstride, in_goff, rows, cols, float(scale),
                                       int(accumulate), _p(out32), ldo32, _p(out16), ldo16, _stream()), "m324_sum_groups")


def cast_transpose_f16(src, N, K, dst, ldo, npad=None, lds=None):
    _chk_f32(src); _chk_f16(dst)
    LAUNCHES[0] += 1
    _l.check(_l.load().m324_cast_transpose_f16(_p(src), lds if lds is not None else K, N, K, _p(dst), ldo, npad if npad is not None else N,
                                               _stream()), "m324_cast_transpose_f16")


def add_block(inp, ld_in, rows, cols, scale, accumulate, out, ldo):
    _chk_f32(inp, out)
    LAUNCHES[0] += 1
    _l.check(_l.load().m324_add_block(_p(inp), ld_in, rows, cols, float(scale), int(accumulate), _p(out), ldo, _stream()), "m324_add_block")


def attn_dot(dO, lddo, O, ldo, rows, H, D, ldd):
    _chk_f16(dO, O); _chk_f32(D)
    LAUNCHES[0] += 1
    _l.check(_l.load().m324_attn_dot(_p(dO), lddo, _p(O), ldo, rows, H, _p(D), ldd, _stream()), "m324_attn_dot")


def attention_bwd(q, k, v, dO, lse, D, dQ, dK, dV, *, B, H, Lq, Lk, q_ld, k_ld, v_ld, do_ld, lse_ld, d_ld, dq_ld, dk_ld, dv_ld, q_rows,
                  kv_rows, q_batch_rows, kv_batch_rows, q_batch_div=1, scale=0.125):
    """Backward of attention(): dQ (accumulated), dK, dV fp32, addressed like q / k / v."""
    _chk_f16(q, k, v, dO)
    _chk_f32(lse, D, dQ, dK, dV)
    a = _l.AttnBwdArgs()
    a.q, a.q_ld, a.q_rows = q.data_ptr(), q_ld, q_rows
    a.k, a.k_ld = k.data_ptr(), k_ld
    a.v, a.v_ld, a.kv_rows = v.data_ptr(), v_ld, kv_rows
    a.B, a.H, a.Lq, a.Lk = B, H, Lq, Lk
    a.q_batch_rows, a.kv_batch_rows, a.q_batch_div = q_batch_rows, kv_batch_rows, q_batch_div
    a.dO, a.do_ld, a.lse, a.lse_ld, a.D, a.d_ld = dO.data_ptr(), do_ld, lse.data_ptr(), lse_ld, D.data_ptr(), d_ld
    a.dQ, a.dq_ld, a.dK, a.dk_ld, a.dV, a.dv_ld = dQ.data_ptr(), dq_ld, dK.data_ptr(), dk_ld, dV.data_ptr(), dv_ld
    a.scale = scale
    LAUNCHES[0] += 1
    _l.check(_l.load().m324_attention_bwd(C.byref(a), _stream()), "m324_attention_bwd")


def scale_by_device_scalars(buf, n, sa, sb=None, cb=0.0):
    """buf[:n] *= (sa + cb * sb) with sa / sb 0-dim CUDA fp32 tensors (no host sync)."""
    _chk_f32(buf, sa, sb)
    LAUNCHES[0] += 1
    _l.check(_l.load().m324_scale_by_device_scalars(_p(buf), n, _p(sa), _p(sb), float(cb), _stream()), "m324_scale_by_device_scalars")


def filter_trajectories(trajs, out, mode, taps=None, mincutoff=1.0, beta=0.007):
    """mode 1: savgol FIR with host `taps` (sequence of floats); mode 2: One-Euro filter."""
    _chk_f32(trajs, out)
    B, T, N, _ = trajs.shape
    arr = (C.c_double * len(taps))(*[float(t) for t in taps]) if taps is not None else None
    LAUNCHES[0] += 1
    _l.check(_l.load().m324_filter_trajectories(_p(trajs), _p(out), B, T, N, int(mode), arr, len(taps) if taps is not None else 0,
                                                float(mincutoff), float(beta), _stream()), "m324_filter_trajectories")


def head3_from_partials(part, groups, b3, rows, out, target, partials):
    """Finish the head fused into a GEMM epilogue (gemm(head_w=, head_part=)); returns the number of MSE partial sums written."""
    _chk_f32(part, b3, out, target, partials)
    n = C.c_int32(0)
    LAUNCHES[0] += 1
    _l.check(_l.load().m324_head3_from_partials(_p(part), int(groups), _p(b3), rows, _p(out), _p(target), _p(partials), C.byref(n), _stream()),
             "m324_head3_from_partials")
    return n.value
