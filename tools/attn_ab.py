"""A/B of the production attention kernels ("v2" below) against the first generation ("v1") at the bench shape (B=32, S=512, 12 heads, d=64):
timing (CUDA events, kernel alone) with and without dropout, plus agreement of their outputs on the same inputs
(b200_set_gemm_debug bit 0x100000 selects the first-generation kernels)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spokennlp_b200 import lib, ops  # noqa: E402
from tools.gemm_sweep import timeit  # noqa: E402

B, S, heads, H = 32, 512, 12, 768
M = B * S
so = lib.load()
torch.manual_seed(0)
qkv = (torch.randn(M, 3 * H, device="cuda") * 1.0).half()
dctx = (torch.randn(M, H, device="cuda") * 0.1).half()
seed = torch.tensor([1234], dtype=torch.int32, device="cuda")
drop = ops.Dropout(seed, 3, 0.1)
mask = torch.ones(B, S, dtype=torch.long, device="cuda")
lens = torch.randint(S // 2, S + 1, (B,), device="cuda")
mask = (torch.arange(S, device="cuda")[None, :] < lens[:, None]).long()
key_bias, kv_len = ops.mask_to_bias(mask)


def run(v1, d, masked):
    so.b200_set_gemm_debug(0x100000 if v1 else 0)
    ctx = torch.zeros(M, H, device="cuda", dtype=torch.float16)
    lse = torch.zeros(B, heads, S, device="cuda")
    dqkv = torch.zeros_like(qkv)
    ws = ops.attn_bwd_workspace(B, heads, S, "cuda")
    kw = dict(key_bias=key_bias, kv_len=kv_len) if masked else {}
    fwd = lambda: ops.attn_fwd(qkv, qkv, ctx, B, heads, S, S, q_col0=0, k_col0=H, v_col0=2 * H, lse2=lse, drop=d, **kw)
    bwd = lambda: ops.attn_bwd(qkv, qkv, dctx, ctx, lse, dqkv, dqkv, ws, B, heads, S, S, q_col0=0, k_col0=H, v_col0=2 * H, dq_col0=0,
                               dk_col0=H, dv_col0=2 * H, drop=d, **kw)
    fwd()
    bwd()
    torch.cuda.synchronize()
    tf, tb = timeit(fwd), timeit(bwd)
    so.b200_set_gemm_debug(0)
    return ctx.float(), lse.clone(), dqkv.float(), tf, tb


fl_f, fl_b = 4.0 * B * S * S * H, 8.0 * B * S * S * H
for masked in (False, True):
    for d in (None, drop):
        c1, l1, g1, tf1, tb1 = run(True, d, masked)
        c2, l2, g2, tf2, tb2 = run(False, d, masked)
        rel = lambda x, y: float((x - y).norm() / (y.norm() + 1e-30))
        fin = lambda l: torch.where(torch.isfinite(l), l, torch.zeros_like(l))
        print(f"masked={masked} drop={d is not None}: fwd v1 {tf1 * 1e6:6.1f} us  v2 {tf2 * 1e6:6.1f} us ({fl_f / tf2 / 1e12:5.0f} TF) | "
              f"bwd(+delta,memset,cast) v1 {tb1 * 1e6:6.1f} us  v2 {tb2 * 1e6:6.1f} us ({fl_b / tb2 / 1e12:5.0f} TF) | "
              f"v2 vs v1: ctx {rel(c2, c1):.2e} lse {rel(fin(l2), fin(l1)):.2e} dqkv {rel(g2, g1):.2e}", flush=True)
