"""Attention kernels alone at the bench shape, for `ncu --set full` captures and for the Sk sweep that separates the
per-item fixed cost from the per-key-block cost:  [B200_ATTN_DROP=0.1] python tools/prof_attn.py [sweep | ln]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spokennlp_b200 import lib, ops  # noqa: E402
from tools.timing import timeit  # noqa: E402

heads, H = 12, 768
so = lib.load()
torch.manual_seed(0)


def make(B, S):
    M = B * S
    qkv = torch.randn(M, 3 * H, device="cuda").half()
    dctx = (torch.randn(M, H, device="cuda") * 0.1).half()
    ctx = torch.zeros(M, H, device="cuda", dtype=torch.float16)
    lse = torch.zeros(B, heads, S, device="cuda")
    dqkv = torch.zeros_like(qkv)
    ws = ops.attn_bwd_workspace(B, heads, S, "cuda")
    p = float(os.environ.get("B200_ATTN_DROP", "0"))       # e.g. 0.1: the training configuration (dropout on the probabilities)
    drop = ops.Dropout(torch.tensor([7], dtype=torch.int32, device="cuda"), 9, p) if p > 0 else None
    fwd = lambda: ops.attn_fwd(qkv, qkv, ctx, B, heads, S, S, q_col0=0, k_col0=H, v_col0=2 * H, lse2=lse, drop=drop)
    bwd = lambda: ops.attn_bwd(qkv, qkv, dctx, ctx, lse, dqkv, dqkv, ws, B, heads, S, S, q_col0=0, k_col0=H, v_col0=2 * H, dq_col0=0,
                               dk_col0=H, dv_col0=2 * H, drop=drop)
    return fwd, bwd


if len(sys.argv) > 1 and sys.argv[1] == "sweep":
    for B, S in ((128, 128), (64, 256), (32, 512), (16, 1024), (8, 2048)):
        fwd, bwd = make(B, S)
        fwd(); bwd()
        tf, tb = timeit(fwd), timeit(bwd)
        fl = 4.0 * B * S * S * H
        print(f"B={B:4d} S={S:5d}: fwd {tf * 1e6:7.1f} us ({fl / tf / 1e12:5.0f} TF)  bwd(+delta,memset,cast) {tb * 1e6:7.1f} us ({2 * fl / tb / 1e12:5.0f} TF)",
              flush=True)
elif len(sys.argv) <= 1:
    fwd, bwd = make(32, 512)
    fwd(); bwd()
    torch.cuda.synchronize()
    print("done")
if len(sys.argv) > 1 and sys.argv[1] == "ln":
    M, Hh = 16384, 768
    dy = (torch.randn(M, Hh, device="cuda") * 0.1).half()
    x = torch.randn(M, Hh, device="cuda")
    mean, rstd = x.mean(1).contiguous(), (x.var(1, unbiased=False) + 1e-12).rsqrt().contiguous()
    gamma = torch.randn(Hh, device="cuda")
    dx = torch.zeros(M, Hh, device="cuda", dtype=torch.float16)
    dg, db, dbias = (torch.zeros(Hh, device="cuda") for _ in range(3))
    ops.layernorm_bwd(dy, x, mean, rstd, gamma, dx, dg, db, dbias=dbias)
    torch.cuda.synchronize()
