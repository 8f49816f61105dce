"""Warp-level timeline of ONE CTA pair of gemm2_kernel (profiling build, see csrc/gemm.cu):

    scripts/build_variant.sh $PWD/scripts/var/timeline.so -DM324_TIMELINE=1
    M324_LIB=scripts/var/timeline.so python scripts/gemm_timeline.py [--shape mlp_up|mlp_down|qkv|fc] [--cluster 10]

Per tile of the traced CTA pair: how long the MMA warp waited for a free accumulator stage (epilogue-bound) vs issued K blocks
(main loop), how long the epilogue warps waited for a full accumulator (MMA / load-bound) vs spent in the epilogue.
"""
import argparse
import ctypes as C
import os
import sys
from collections import defaultdict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motion324_b200 import lib as _l, ops  # noqa: E402

EV = {1: "load_tile", 8: "acc_wait", 9: "acc_ok", 10: "mma_issued", 16: "epi_wait", 17: "epi_start", 18: "epi_end"}
SHAPES = {"mlp_up": (10368, 3072, 768, dict(act=1)), "mlp_down": (10368, 768, 3072, dict(resid=True)),
          "qkv": (10368, 2304, 768, dict(qk=True)), "fc": (10368, 768, 768, dict(resid=True))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="mlp_up", choices=list(SHAPES))
    ap.add_argument("--cluster", type=int, default=10)
    a = ap.parse_args()
    lib = _l.load()
    if not hasattr(lib, "m324_timeline_set_gemm"):
        raise SystemExit("this libm324 was built without -DM324_TIMELINE=1 (set M324_LIB to the profiling build)")
    lib.m324_timeline_set_gemm.argtypes = [C.c_void_p, C.c_int, C.c_int]
    M, N, K, opt = SHAPES[a.shape]
    A = torch.randn(M, K, device="cuda").half()
    W = (torch.randn(N, K, device="cuda") * 0.02).half()
    out16 = torch.empty(M, N, device="cuda", dtype=torch.float16)
    x = torch.zeros(M, N, device="cuda")
    qn = torch.ones(64, device="cuda")

    def run():
        if opt.get("resid"):
            ops.gemm(A, W, M, N, K, resid=x, ldr=N, out32=x, ldo32=N)
        elif opt.get("qk"):
            ops.gemm(A, W, M, N, K, out16=out16, ldo16=N, qn_w=qn, kn_w=qn, qk_eps=1e-5, qk_cols=768)
        else:
            ops.gemm(A, W, M, N, K, act=opt.get("act", 0), out16=out16, ldo16=N)
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    cap = 1 << 16
    buf = torch.zeros(cap, dtype=torch.int64, device="cuda")
    assert lib.m324_timeline_set_gemm(C.c_void_p(buf.data_ptr()), cap, a.cluster) == 0
    run()
    torch.cuda.synchronize()
    lib.m324_timeline_set_gemm(None, 0, -1)
    raw = buf.cpu().numpy().astype("uint64")
    n = int(raw[0])
    recs = sorted(((int(r) >> 16, (int(r) >> 8) & 0xFF, int(r) & 0xFF) for r in raw[1:1 + min(n, cap - 1)]))
    if not recs:
        raise SystemExit("no records: is --cluster inside the grid (74 CTA pairs)?")
    t0 = recs[0][0]
    per_warp = defaultdict(list)
    for t, w, e in recs:
        per_warp[w].append((t - t0, EV.get(e, str(e))))
    print(f"{a.shape}: M={M} N={N} K={K}; {n} records from CTA pair {a.cluster}; span {recs[-1][0] - t0} clk; warps (rank * 16 + warp) {sorted(per_warp)}")
    for w in sorted(per_warp):
        ev = per_warp[w]
        role = {0: "TMA producer", 1: "MMA issuer"}.get(w % 16, "epilogue")
        if role == "MMA issuer":
            waits = [b[0] - a_[0] for a_, b in zip(ev, ev[1:]) if a_[1] == "acc_wait" and b[1] == "acc_ok"]
            loops = [b[0] - a_[0] for a_, b in zip(ev, ev[1:]) if a_[1] == "acc_ok" and b[1] == "mma_issued"]
            print(f"warp {w:2d} {role}: tiles {len(loops)}; wait for a free accumulator avg {sum(waits) / max(len(waits), 1):.0f} clk, "
                  f"issue main loop avg {sum(loops) / max(len(loops), 1):.0f} clk; per tile: " + " ".join(f"{x}+{y}" for x, y in zip(waits, loops)))
        elif role == "epilogue":
            waits = [b[0] - a_[0] for a_, b in zip(ev, ev[1:]) if a_[1] == "epi_wait" and b[1] == "epi_start"]
            epis = [b[0] - a_[0] for a_, b in zip(ev, ev[1:]) if a_[1] == "epi_start" and b[1] == "epi_end"]
            print(f"warp {w:2d} {role}: wait for a full accumulator avg {sum(waits) / max(len(waits), 1):.0f} clk, epilogue avg "
                  f"{sum(epis) / max(len(epis), 1):.0f} clk; per tile: " + " ".join(f"{x}+{y}" for x, y in zip(waits, epis)))
        else:
            print(f"warp {w:2d} {role}: tile starts at " + " ".join(str(t) for t, e in ev if e == "load_tile"))


if __name__ == "__main__":
    main()
