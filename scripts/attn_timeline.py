"""Warp-level timeline of ONE CTA of the attention forward kernel (global-layer shape), from a profiling build:

    scripts/build_variant.sh $PWD/scripts/var/timeline.so -DM324_TIMELINE=1
    M324_LIB=scripts/var/timeline.so python scripts/attn_timeline.py [--frames 32] [--cta 200] [--steps 20:24]

Lane 0 of every warp of the chosen CTA records (clock64, warp, event) at the points marked TL(...) in csrc/attention.cu.
Prints, per K/V step of the requested window, when each softmax warp of SM sub-partition 0 (warps 4 and 8: Q tile 0 / 1)
waited for S, finished loading S, took its row max, got its P V back, was given the MUFU turn, finished its exponentials
and handed P over, and when the two MMA-issuing warps issued Q K^T / P V -- plus the average time per phase.  This is the
view ncu's sampler cannot give: WHERE the MUFU unit idles between the two warps that share it.
"""
import argparse
import ctypes as C
import os
import sys
from collections import defaultdict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motion324_b200 import lib as _l, ops  # noqa: E402

EV_BWD = {1: "top", 2: "s_ready", 3: "exp_done", 4: "mma_done_ok", 5: "pt_stored", 6: "dq_out", 7: "dp_loaded", 8: "p_ready",
          16: "s_wait", 17: "s_issued", 18: "g_wait", 19: "g_issued"}
EV = {1: "s_wait", 2: "s_ready", 3: "s_loaded", 4: "max_done", 5: "odone_ok", 6: "turn_ok", 7: "exp_done", 8: "turn_passed",
      9: "p_arrived", 16: "qk_wait", 17: "qk_issued", 18: "pv_wait", 19: "pv_issued"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=32)
    ap.add_argument("--cta", type=int, default=200)
    ap.add_argument("--steps", default="20:24", help="K/V steps (forward) / query tiles (backward) to print, first:last")
    ap.add_argument("--bwd", action="store_true", help="trace attn_bwd_kernel instead (CTA = one 128-key tile walking the query tiles)")
    a = ap.parse_args()
    lib = _l.load()
    if a.bwd:
        return main_bwd(a, lib)
    if not hasattr(lib, "m324_timeline_set"):
        raise SystemExit("this libm324 was built without -DM324_TIMELINE=1 (set M324_LIB to the profiling build)")
    lib.m324_timeline_set.argtypes = [C.c_void_p, C.c_int, C.c_int]
    d, H, L = 768, 12, a.frames * 324
    qkv = torch.randn(L, 3 * d, device="cuda").half()
    o = torch.empty(L, d, device="cuda", dtype=torch.float16)
    kw = dict(B=1, H=H, Lq=L, Lk=L, q_ld=3 * d, k_ld=3 * d, v_ld=3 * d, o_ld=d, q_rows=L, kv_rows=L, q_batch_rows=L, kv_batch_rows=L, scale=0.125)
    cap = 1 << 18
    buf = torch.zeros(cap, dtype=torch.int64, device="cuda")
    for _ in range(2):
        ops.attention(qkv, qkv[:, d:], qkv[:, 2 * d:], o, **kw)
    torch.cuda.synchronize()
    assert lib.m324_timeline_set(C.c_void_p(buf.data_ptr()), cap, a.cta) == 0
    ops.attention(qkv, qkv[:, d:], qkv[:, 2 * d:], o, **kw)
    torch.cuda.synchronize()
    lib.m324_timeline_set(None, 0, -1)
    raw = buf.cpu().numpy().astype("uint64")
    n = int(raw[0])
    recs = sorted(((int(r) >> 16, (int(r) >> 8) & 0xFF, int(r) & 0xFF) for r in raw[1:1 + min(n, cap - 1)]))
    if not recs:
        raise SystemExit("no records: is --cta inside the grid?")
    t0 = recs[0][0]
    per_warp = defaultdict(list)
    for t, w, e in recs:
        per_warp[w].append((t - t0, EV.get(e, str(e))))
    print(f"{n} records from CTA {a.cta}; warps {sorted(per_warp)}; span {recs[-1][0] - t0} clk")
    # K/V step index of a softmax warp = number of 's_wait' records seen so far
    first, last = (int(x) for x in a.steps.split(":"))
    for w in sorted(per_warp):
        step, rows = -1, defaultdict(dict)
        key = "s_wait" if w >= 4 else "qk_wait"
        for t, e in per_warp[w]:
            if e == key:
                step += 1
            rows[step][e] = t
        print(f"\nwarp {w} ({'softmax Q tile %d' % ((w - 4) >> 2) if w >= 4 else 'MMA issue Q tile %d' % (w - 1)}):")
        for s_ in range(first, last + 1):
            if s_ in rows:
                base = rows[s_].get(key, 0)
                print(f"  step {s_:3d} @ {base:8d}: " + "  ".join(f"{e}+{t - base}" for e, t in sorted(rows[s_].items(), key=lambda kv: kv[1]) if e != key))
        # phase averages over the steady state (steps 4 .. n-4)
        steps_ = [s_ for s_ in rows if 4 <= s_ < step - 4]
        if w >= 4 and steps_:
            def avg(a_, b_):
                v = [rows[s_][b_] - rows[s_][a_] for s_ in steps_ if a_ in rows[s_] and b_ in rows[s_]]
                return sum(v) / len(v) if v else float("nan")
            period = [rows[s_ + 1]["s_wait"] - rows[s_]["s_wait"] for s_ in steps_ if s_ + 1 in rows and "s_wait" in rows[s_ + 1]]
            print(f"  avg clk/step {sum(period) / len(period):.0f}: wait S {avg('s_wait', 's_ready'):.0f} | load S {avg('s_ready', 's_loaded'):.0f} | max "
                  f"{avg('s_loaded', 'max_done'):.0f} | wait PV {avg('max_done', 'odone_ok'):.0f} | wait turn {avg('odone_ok', 'turn_ok'):.0f} | exp "
                  f"{avg('turn_ok', 'exp_done'):.0f} | store P + arrive {avg('exp_done', 'p_arrived'):.0f}")


def main_bwd(a, lib):
    if not hasattr(lib, "m324_timeline_set_bwd"):
        raise SystemExit("this libm324 was built without -DM324_TIMELINE=1 (set M324_LIB to the profiling build)")
    lib.m324_timeline_set_bwd.argtypes = [C.c_void_p, C.c_int, C.c_int]
    d, H, L = 768, 12, a.frames * 324
    qkv = torch.randn(L, 3 * d, device="cuda").half()
    o = torch.empty(L, d, device="cuda", dtype=torch.float16)
    lse = torch.empty(L, H, device="cuda")
    kw = dict(B=1, H=H, Lq=L, Lk=L, q_ld=3 * d, k_ld=3 * d, v_ld=3 * d, q_rows=L, kv_rows=L, q_batch_rows=L, kv_batch_rows=L, scale=0.125)
    ops.attention(qkv, qkv[:, d:], qkv[:, 2 * d:], o, o_ld=d, lse=lse, lse_ld=H, **kw)
    dO = (torch.randn(L, d, device="cuda") * 0.01).half()
    D = torch.empty(L, H, device="cuda")
    ops.attn_dot(dO, d, o, d, L, H, D, H)
    dqkv = torch.zeros(L, 3 * d, device="cuda")
    run = lambda: ops.attention_bwd(qkv, qkv[:, d:], qkv[:, 2 * d:], dO, lse, D, dqkv, dqkv[:, d:], dqkv[:, 2 * d:], do_ld=d, lse_ld=H, d_ld=H,
                                    dq_ld=3 * d, dk_ld=3 * d, dv_ld=3 * d, **kw)
    run()
    torch.cuda.synchronize()
    cap = 1 << 18
    buf = torch.zeros(cap, dtype=torch.int64, device="cuda")
    assert lib.m324_timeline_set_bwd(C.c_void_p(buf.data_ptr()), cap, a.cta) == 0
    run()
    torch.cuda.synchronize()
    lib.m324_timeline_set_bwd(None, 0, -1)
    raw = buf.cpu().numpy().astype("uint64")
    n = int(raw[0])
    recs = sorted(((int(r) >> 16, (int(r) >> 8) & 0xFF, int(r) & 0xFF) for r in raw[1:1 + min(n, cap - 1)]))
    if not recs:
        raise SystemExit("no records: is --cta inside the grid?")
    t0 = recs[0][0]
    per_warp = defaultdict(list)
    for t, w, e in recs:
        per_warp[w].append((t - t0, EV_BWD.get(e, str(e))))
    print(f"{n} records from CTA {a.cta}; warps {sorted(per_warp)}; span {recs[-1][0] - t0} clk")
    first, last = (int(x) for x in a.steps.split(":"))
    for w in sorted(per_warp):
        key = "top" if w >= 4 else "g_wait"
        step, rows = -1, defaultdict(dict)
        for t, e in per_warp[w]:
            if e == key:
                step += 1
            rows[step].setdefault(e, t)
        print(f"\nwarp {w} ({'P^T / dS^T threads' if w >= 4 else 'MMA issue'}):")
        for s_ in range(first, last + 1):
            if s_ in rows:
                base = rows[s_].get(key, 0)
                print(f"  tile {s_:3d} @ {base:8d}: " + "  ".join(f"{e}+{t - base}" for e, t in sorted(rows[s_].items(), key=lambda kv: kv[1]) if e != key))
        steps_ = [s_ for s_ in rows if 2 <= s_ < step - 2]
        if w >= 4 and steps_:
            def avg(a_, b_):
                v = [rows[s_][b_] - rows[s_][a_] for s_ in steps_ if a_ in rows[s_] and b_ in rows[s_]]
                return sum(v) / len(v) if v else float("nan")
            period = [rows[s_ + 1]["top"] - rows[s_]["top"] for s_ in steps_ if s_ + 1 in rows and "top" in rows[s_ + 1]]
            print(f"  avg clk/tile {sum(period) / len(period):.0f}: wait S/dP {avg('top', 's_ready'):.0f} | exp {avg('s_ready', 'exp_done'):.0f} | wait MMAs "
                  f"{avg('exp_done', 'mma_done_ok'):.0f} | store P^T {avg('mma_done_ok', 'pt_stored'):.0f} | dQ out {avg('pt_stored', 'dq_out'):.0f} | load dP "
                  f"{avg('dq_out', 'dp_loaded'):.0f} | dS + arrive {avg('dp_loaded', 'p_ready'):.0f}")


if __name__ == "__main__":
    main()
