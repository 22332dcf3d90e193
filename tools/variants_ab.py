"""A/B of the opt-in kernel variants against the round-1 kernels they replace, on the bench shapes.
Each kernel is timed alone with CUDA events (operands >> L2); one JSON line per comparison.

    python tools/variants_ab.py            (on a B200)
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spokennlp_b200 import lib, ops  # noqa: E402


def timeit(fn, iters=30):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def main():
    M, H, I, B, S, heads = 16384, 768, 3072, 32, 512, 12
    dev, f16 = "cuda", torch.float16
    seed = torch.tensor([7], dtype=torch.int32, device=dev)
    for name, K in (("out_proj  16384x768x768 ", H), ("ffn_down  16384x768x3072", I)):
        a = torch.randn(M, K, device=dev, dtype=f16)
        w = torch.randn(H, K, device=dev, dtype=f16) * 0.02
        bias = torch.zeros(H, device=dev)
        res = torch.randn(M, H, device=dev)
        out = torch.empty(M, H, device=dev)
        acc = res.clone()
        fl = 2.0 * M * H * K / 1e12
        for p in (0.0, 0.1):
            drop = ops.Dropout(seed, 5, p) if p > 0 else None
            t0 = timeit(lambda: ops.gemm(a, w, out, epilogue=ops.EPI_BIAS_RES32, bias=bias, aux=res, drop=drop))
            t1 = timeit(lambda: ops.gemm_resadd(a, w, acc, bias, drop=drop))
            t2 = timeit(lambda: ops.gemm_resadd(a, w, acc, bias, drop=drop, stream_k=True))
            print(json.dumps({"shape": name.strip(), "dropout": p, "res32_us": t0 * 1e6, "resadd_us": t1 * 1e6, "streamk_us": t2 * 1e6,
                              "res32_tflops": fl / t0, "resadd_tflops": fl / t1, "streamk_tflops": fl / t2}), flush=True)
    # output-projection dgrad, with and without the fused row statistic (+ the separate kernel it replaces: measured inside attn_bwd)
    dy = torch.randn(M, H, device=dev, dtype=f16)
    w = torch.randn(H, H, device=dev, dtype=f16) * 0.02
    ctx = torch.randn(M, H, device=dev, dtype=f16)
    dctx = torch.empty(M, H, device=dev, dtype=f16)
    ws = ops.attn_bwd_workspace(B, heads, S, dev)
    t0 = timeit(lambda: ops.gemm(dy, w, dctx, b_layout=1))
    t1 = timeit(lambda: ops.gemm_dgrad_delta(dy, w, ctx, dctx, ws, B, heads, S))
    qkv = torch.randn(M, 3 * H, device=dev, dtype=f16)
    lse = torch.zeros(B, heads, S, device=dev)
    dqkv = torch.empty_like(qkv)
    kw = dict(q_col0=0, k_col0=H, v_col0=2 * H, dq_col0=0, dk_col0=H, dv_col0=2 * H)
    t2 = timeit(lambda: ops.attn_bwd(qkv, qkv, dctx, ctx, lse, dqkv, dqkv, ws, B, heads, S, S, **kw))
    t3 = timeit(lambda: ops.attn_bwd(qkv, qkv, dctx, ctx, lse, dqkv, dqkv, ws, B, heads, S, S, delta_ready=True, **kw))
    print(json.dumps({"shape": "out_proj dgrad 16384x768x768 + attn_bwd", "dgrad_us": t0 * 1e6, "dgrad_delta_us": t1 * 1e6,
                      "attn_bwd_us": t2 * 1e6, "attn_bwd_delta_ready_us": t3 * 1e6,
                      "pair_before_us": (t0 + t2) * 1e6, "pair_after_us": (t1 + t3) * 1e6}), flush=True)
    # persistent attention kernels: per-thread vs warp-elected mbarrier arrivals, with and without dropout
    ctx2 = torch.empty(M, H, device=dev, dtype=f16)
    lse2 = torch.empty(B, heads, S, device=dev)
    for p in (0.0, 0.1):
        drop = ops.Dropout(seed, 9, p) if p > 0 else None
        row = {"shape": "attention B32 S512 h12", "dropout": p}
        for variant, tag in ((0, "default"), (1, "elect"), (3, "elect_wait"), (4, "bwd16"), (8, "fwd16")):
            lib.load().b200_set_attn_variant(variant)
            row[f"fwd_{tag}_us"] = 1e6 * timeit(lambda: ops.attn_fwd(qkv, qkv, ctx2, B, heads, S, S, q_col0=0, k_col0=H, v_col0=2 * H, lse2=lse2, drop=drop))
            row[f"bwd_{tag}_us"] = 1e6 * timeit(lambda: ops.attn_bwd(qkv, qkv, dctx, ctx2, lse2, dqkv, dqkv, ws, B, heads, S, S, drop=drop, **kw))
        lib.load().b200_set_attn_variant(0)
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
