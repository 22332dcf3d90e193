"""Attention kernels at a fixed token count and growing sequence length: separates the per-item cost (prologue, finalise,
context store) from the per-block cost of the persistent kernels.  items = B * heads * S / 256 (forward) is constant.

    python tools/attn_scaling.py            (on a B200)
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spokennlp_b200 import ops  # noqa: E402
from tools.timing import timeit  # noqa: E402


def main():
    H, heads = 768, 12
    dev, f16 = "cuda", torch.float16
    seed = torch.tensor([7], dtype=torch.int32, device=dev)
    for B, S in ((32, 512), (16, 1024), (8, 2048), (4, 4096), (64, 256), (128, 128)):
        M = B * S
        qkv = torch.randn(M, 3 * H, device=dev, dtype=f16)
        dctx = torch.randn(M, H, device=dev, dtype=f16)
        dqkv = torch.empty(M, 3 * H, device=dev, dtype=f16)
        ctx = torch.empty(M, H, device=dev, dtype=f16)
        lse = torch.empty(B, heads, S, device=dev)
        ws = ops.attn_bwd_workspace(B, heads, S, dev)
        kw = dict(q_col0=0, k_col0=H, v_col0=2 * H, dq_col0=0, dk_col0=H, dv_col0=2 * H)
        for p in (0.0, 0.1):
            drop = ops.Dropout(seed, 9, p) if p > 0 else None
            row = {"B": B, "S": S, "dropout": p, "blocks_per_item": S // 128}
            row["fwd_us"] = 1e6 * timeit(lambda: ops.attn_fwd(qkv, qkv, ctx, B, heads, S, S, q_col0=0, k_col0=H, v_col0=2 * H, lse2=lse, drop=drop))
            row["bwd_us"] = 1e6 * timeit(lambda: ops.attn_bwd(qkv, qkv, dctx, ctx, lse, dqkv, dqkv, ws, B, heads, S, S, drop=drop, delta_ready=True, dq_half=True, **kw))
            fl = 4.0 * B * heads * S * S * 64
            row["fwd_tflops"] = fl / row["fwd_us"] / 1e6
            row["bwd_tflops"] = 2.5 * fl / row["bwd_us"] / 1e6
            print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
