"""Timeline of CTA 0 of the persistent attention kernels from a -DB200_ATT_TRACE build of the library (ptx.cuh: ATT_TRACE).

    nvcc -gencode arch=compute_100a,code=sm_100a -DB200_SRC_HASH=\\"unknown\\" -DB200_ATT_TRACE -O3 -std=c++17 -Xcompiler -fPIC \\
         --expt-relaxed-constexpr -shared -o tools/micro/libb200enc_trace.so spokennlp_b200/csrc/api.cu -lcudart_static -lrt -ldl -lpthread
    python tools/attn_trace.py [fwd|bwd] [dropout]          (on a B200)

Prints, per warp, the clock of every event relative to the kernel's first record, and a per-phase summary: where a block's
period goes (waiting for the tensor core, pulling scores, softmax math, waiting for the previous product, publishing) and what
an item boundary costs.
"""
import ctypes as C
import os
import sys
from collections import defaultdict

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spokennlp_b200 import lib  # noqa: E402

lib.LIB_PATH = os.path.join(ROOT, "tools", "micro", "libb200enc_trace.so")
lib.is_stale = lambda: False
from spokennlp_b200 import ops  # noqa: E402


def read_trace():
    L = lib.load()
    L.b200_att_trace_read.argtypes = [C.c_void_p, C.c_void_p]
    L.b200_att_trace_read.restype = C.c_int
    buf = np.zeros((16, 4096), dtype=np.uint64)
    cnt = np.zeros(16, dtype=np.uint32)
    L.b200_att_trace_read(buf.ctypes.data, cnt.ctypes.data)
    warp, ev, clk = [], [], []
    for w in range(16):
        rec = buf[w, :min(int(cnt[w]), 4096)]
        warp += [w] * len(rec)
        ev += ((rec >> np.uint64(48)) & np.uint64(0xFF)).astype(np.int64).tolist()
        clk += (rec & np.uint64(0xFFFFFFFFFFFF)).astype(np.int64).tolist()
    return np.array(warp), np.array(ev), np.array(clk)


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
    p = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1
    B, S, heads, H = 32, 512, 12, 768
    dev, f16 = "cuda", torch.float16
    M = B * S
    seed = torch.tensor([7], dtype=torch.int32, device=dev)
    qkv = torch.randn(M, 3 * H, device=dev, dtype=f16)
    dctx = torch.randn(M, H, device=dev, dtype=f16)
    dqkv = torch.zeros(M, 3 * H, device=dev, dtype=f16)
    ctx = torch.empty(M, H, device=dev, dtype=f16)
    lse = torch.empty(B, heads, S, device=dev)
    ws = ops.attn_bwd_workspace(B, heads, S, dev)
    drop = ops.Dropout(seed, 9, p) if p > 0 else None
    kw = dict(q_col0=0, k_col0=H, v_col0=2 * H, dq_col0=0, dk_col0=H, dv_col0=2 * H)
    fwd = lambda: ops.attn_fwd(qkv, qkv, ctx, B, heads, S, S, q_col0=0, k_col0=H, v_col0=2 * H, lse2=lse, drop=drop)
    bwd = lambda: ops.attn_bwd(qkv, qkv, dctx, ctx, lse, dqkv, dqkv, ws, B, heads, S, S, drop=drop, dq_half=True, **kw)
    for _ in range(3):
        fwd()
        bwd()
    read_trace()                                   # drop the warm-up records
    (fwd if which == "fwd" else bwd)()
    warp, ev, clk = read_trace()
    t0 = clk.min()
    clk = clk - t0
    print(f"{which} dropout {p}: {len(clk)} records, CTA 0 span {clk.max()} clk")
    per_warp = defaultdict(list)
    for w, e, c in sorted(zip(warp.tolist(), ev.tolist(), clk.tolist()), key=lambda r: (r[0], r[2])):
        per_warp[w].append((e, c))
    for w in sorted(per_warp):
        print(f"warp {w:2d}: " + " ".join(f"{e}@{c}" for e, c in per_warp[w][:80]))
    # phase durations per softmax warp: consecutive event pairs (e_prev -> e_next)
    print("\nmean clocks between consecutive events of a warp (event pair: count, mean, max), softmax warps only:")
    agg = defaultdict(list)
    for w, recs in per_warp.items():
        if any(e >= 10 for e, _ in recs):
            continue
        for (e0, c0), (e1, c1) in zip(recs, recs[1:]):
            agg[(e0, e1)].append(c1 - c0)
    tot = sum(sum(v) for v in agg.values())
    for k in sorted(agg, key=lambda k: -sum(agg[k])):
        v = agg[k]
        print(f"  {k[0]:2d} -> {k[1]:2d}: n={len(v):4d} mean={sum(v) / len(v):8.1f} max={max(v):6d}  share={sum(v) / tot:6.1%}")


if __name__ == "__main__":
    main()
