"""Hot kernels timed alone on the bench shapes with CUDA events (operands >> L2), incl. each fused epilogue against the
unfused pair it replaces; one JSON line per comparison.

    python tools/kernel_timings.py            (on a B200)
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spokennlp_b200 import lib, ops  # noqa: E402


from tools.timing import timeit  # noqa: E402


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
            print(json.dumps({"shape": name.strip(), "dropout": p, "res32_us": t0 * 1e6, "resadd_us": t1 * 1e6,
                              "res32_tflops": fl / t0, "resadd_tflops": fl / t1}), flush=True)
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
    # persistent attention kernels, with and without dropout
    ctx2 = torch.empty(M, H, device=dev, dtype=f16)
    lse2 = torch.empty(B, heads, S, device=dev)
    for p in (0.0, 0.1):
        drop = ops.Dropout(seed, 9, p) if p > 0 else None
        row = {"shape": "attention B32 S512 h12", "dropout": p}
        row["fwd_us"] = 1e6 * timeit(lambda: ops.attn_fwd(qkv, qkv, ctx2, B, heads, S, S, q_col0=0, k_col0=H, v_col0=2 * H, lse2=lse2, drop=drop))
        row["bwd_us"] = 1e6 * timeit(lambda: ops.attn_bwd(qkv, qkv, dctx, ctx2, lse2, dqkv, dqkv, ws, B, heads, S, S, drop=drop, **kw))
        row["bwd_delta_ready_us"] = 1e6 * timeit(lambda: ops.attn_bwd(qkv, qkv, dctx, ctx2, lse2, dqkv, dqkv, ws, B, heads, S, S, drop=drop, delta_ready=True, **kw))
        print(json.dumps(row), flush=True)
    # FFN-down dgrad: plain dGELU epilogue + separate column sum vs the fused one; FFN-up with the GELU epilogue
    dy = torch.randn(M, H, device=dev, dtype=f16)
    w2 = torch.randn(H, I, device=dev, dtype=f16) * 0.02
    dact = torch.rand(M, I, device=dev).half()
    dz = torch.empty(M, I, device=dev, dtype=f16)
    db = torch.zeros(I, device=dev)
    one = torch.ones(1, device=dev)
    t0 = timeit(lambda: ops.gemm(dy, w2, dz, b_layout=1, epilogue=ops.EPI_DGELU, aux=dact))
    t1 = timeit(lambda: ops.colsum(dz, db, one))
    t2 = timeit(lambda: ops.gemm_dgelu_colsum(dy, w2, dact, dz, db, one))
    print(json.dumps({"shape": "ffn_down dgrad 16384x3072x768", "dgelu_us": t0 * 1e6, "colsum_us": t1 * 1e6, "dgelu_colsum_us": t2 * 1e6,
                      "fused_tflops": 2.0 * M * I * H / 1e12 / t2}), flush=True)
    x = torch.randn(M, H, device=dev, dtype=f16)
    w1 = torch.randn(I, H, device=dev, dtype=f16) * 0.02
    b1 = torch.zeros(I, device=dev)
    hbuf, dbuf = torch.empty(M, I, device=dev, dtype=f16), torch.empty(M, I, device=dev, dtype=f16)
    t0 = timeit(lambda: ops.gemm(x, w1, hbuf, epilogue=ops.EPI_BIAS_GELU, bias=b1, out2=dbuf))
    t1 = timeit(lambda: ops.gemm(x, w1, hbuf, epilogue=ops.EPI_BIAS, bias=b1))
    print(json.dumps({"shape": "ffn_up 16384x3072x768", "bias_gelu_us": t0 * 1e6, "bias_only_us": t1 * 1e6,
                      "bias_gelu_tflops": 2.0 * M * I * H / 1e12 / t0, "bias_only_tflops": 2.0 * M * I * H / 1e12 / t1}), flush=True)
    # BASELINE config 4's bf16 arm: the mmvts projector contraction (visual features, K = 3328) on bf16 against fp16 operands
    Mv, Kv = 16 * 300, 3328
    xv = torch.randn(Mv, Kv, device=dev)
    wv = torch.randn(H, Kv, device=dev) * 0.02
    ov = torch.empty(Mv, H, device=dev)
    bv = torch.zeros(H, device=dev)
    x16v, w16v = xv.half(), wv.half()
    xbv, wbv = xv.to(torch.bfloat16), wv.to(torch.bfloat16)
    t0 = timeit(lambda: ops.gemm(x16v, w16v, ov, epilogue=ops.EPI_BIAS, bias=bv))
    e16 = float((ov.double() - (xv.double() @ wv.double().t())).norm() / (xv.double() @ wv.double().t()).norm())
    t1 = timeit(lambda: ops.gemm_bf16(xbv, wbv, ov, epilogue=ops.EPI_BIAS, bias=bv))
    eb = float((ov.double() - (xv.double() @ wv.double().t())).norm() / (xv.double() @ wv.double().t()).norm())
    print(json.dumps({"shape": f"mmvts projector {Mv}x{H}x{Kv}", "fp16_us": t0 * 1e6, "bf16_us": t1 * 1e6, "fp16_rel_err_vs_fp32": e16,
                      "bf16_rel_err_vs_fp32": eb, "tflops_bf16": 2.0 * Mv * H * Kv / 1e12 / t1}), flush=True)
    # token-classification head (HBM-bound: 25 MB of fp16 activations per launch)
    W = torch.randn(2, H, device=dev) * 0.02
    bb = torch.zeros(2, device=dev)
    logits = torch.empty(M, 2, device=dev)
    lc = lib.load()
    import ctypes as C
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    call = lambda: lc.b200_cls_head_fwd(C.c_void_p(x.data_ptr()), C.c_void_p(W.data_ptr()), C.c_void_p(bb.data_ptr()), C.c_void_p(logits.data_ptr()), None, M, H, 2, st)
    t0 = timeit(call, iters=50)
    print(json.dumps({"shape": "cls_head_fwd 16384x768 -> 2", "us": t0 * 1e6, "GBps": M * H * 2 / t0 / 1e9}), flush=True)


if __name__ == "__main__":
    main()
