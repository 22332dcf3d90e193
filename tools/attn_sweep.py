"""Attention kernels at the bench shape under the measurement knobs (b200_set_gemm_debug bits 0x10000.. apply to attn_bwd)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spokennlp_b200 import lib, ops  # noqa: E402
from tools.gemm_sweep import timeit  # noqa: E402

B, S, heads, H = 32, 512, 12, 768
M = B * S
so = lib.load()
qkv = torch.randn(M, 3 * H, device="cuda", dtype=torch.float16)
ctx = torch.empty(M, H, device="cuda", dtype=torch.float16)
lse = torch.empty(B, heads, S, device="cuda")
dqkv = torch.empty_like(qkv)
ws = ops.attn_bwd_workspace(B, heads, S, "cuda")
fwd = lambda: ops.attn_fwd(qkv, qkv, ctx, B, heads, S, S, q_col0=0, k_col0=H, v_col0=2 * H, lse2=lse)
bwd = lambda: ops.attn_bwd(qkv, qkv, ctx, ctx, lse, dqkv, dqkv, ws, B, heads, S, S, q_col0=0, k_col0=H, v_col0=2 * H, dq_col0=0, dk_col0=H,
                           dv_col0=2 * H)
fwd()
print(f"attn_fwd {timeit(fwd) * 1e6:7.1f} us")
for dbg in (0, 0x10000, 0x20000, 0x40000, 0x30000, 0x50000, 0x60000, 0x70000):
    so.b200_set_gemm_debug(dbg)
    print(f"attn_bwd (+delta, memset, cast) dbg={dbg:#x}: {timeit(bwd) * 1e6:7.1f} us", flush=True)
so.b200_set_gemm_debug(0)
