"""Launch each hot kernel a few times at the bench shapes (B=32, S=512, BERT-base) so that ncu can capture them:

    ncu --set full --clock-control none --import-source on -o gpurun_out/prof_r01 python tools/prof_kernels.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spokennlp_b200 import ops  # noqa: E402

B, S, H, I, heads = 32, 512, 768, 3072, 12
M = B * S
dev, f16 = "cuda", torch.float16
reps = int(os.environ.get("REPS", "2"))
torch.manual_seed(0)
x = torch.randn(M, H, device=dev, dtype=f16)
wqkv = torch.randn(3 * H, H, device=dev, dtype=f16) * 0.02
w1 = torch.randn(I, H, device=dev, dtype=f16) * 0.02
w2 = torch.randn(H, I, device=dev, dtype=f16) * 0.02
wo = torch.randn(H, H, device=dev, dtype=f16) * 0.02
bq, b1, bo = torch.zeros(3 * H, device=dev), torch.zeros(I, device=dev), torch.zeros(H, device=dev)
qkv = torch.empty(M, 3 * H, device=dev, dtype=f16)
h = torch.empty(M, I, device=dev, dtype=f16)
z = torch.empty(M, I, device=dev, dtype=f16)
pre = torch.empty(M, H, device=dev, dtype=torch.float32)
x32 = torch.randn(M, H, device=dev, dtype=torch.float32)
ctx = torch.empty(M, H, device=dev, dtype=f16)
lse = torch.empty(B, heads, S, device=dev)
g, b = torch.ones(H, device=dev), torch.zeros(H, device=dev)
mean, rstd = torch.empty(M, device=dev), torch.empty(M, device=dev)
y = torch.empty(M, H, device=dev, dtype=f16)
dx = torch.empty(M, H, device=dev, dtype=f16)
dz = torch.empty(M, I, device=dev, dtype=f16)
dqkv = torch.empty(M, 3 * H, device=dev, dtype=f16)
gw1, gw2, gwo, gwqkv = (torch.zeros(I, H, device=dev), torch.zeros(H, I, device=dev), torch.zeros(H, H, device=dev),
                        torch.zeros(3 * H, H, device=dev))
dg, db_, dbias = torch.zeros(H, device=dev), torch.zeros(H, device=dev), torch.zeros(H, device=dev)
ws = ops.attn_bwd_workspace(B, heads, S, dev)

for _ in range(reps):
    ops.gemm(x, wqkv, qkv, epilogue=ops.EPI_BIAS, bias=bq)                                   # QKV projection
    ops.attn_fwd(qkv, qkv, ctx, B, heads, S, S, q_col0=0, k_col0=H, v_col0=2 * H, lse2=lse)   # attention fwd
    ops.gemm(ctx, wo, pre, epilogue=ops.EPI_BIAS_RES32, bias=bo, aux=x32)                      # out-proj + residual
    ops.layernorm_fwd(pre, g, b, 1e-12, y=y, mean=mean, rstd=rstd)                           # LN fwd
    ops.gemm(y, w1, h, epilogue=ops.EPI_BIAS_GELU, bias=b1, out2=z)                          # FFN up + GELU
    ops.gemm(h, w2, pre, epilogue=ops.EPI_BIAS_RES32, bias=bo, aux=x32)                        # FFN down + residual
    ops.layernorm_bwd(y, pre, mean, rstd, g, dx, dg, db_, dbias=dbias)                       # LN bwd
    ops.gemm(dx, h, gw2, a_layout=1, b_layout=1, epilogue=ops.EPI_ATOMIC, k_splits=ops.wgrad_splits(H, I, M))    # wgrad FFN down
    ops.gemm(dx, w2, dz, b_layout=1, epilogue=ops.EPI_DGELU, aux=z)                          # dgrad FFN down + dGELU
    ops.colsum(dz, b1)
    ops.gemm(dz, y, gw1, a_layout=1, b_layout=1, epilogue=ops.EPI_ATOMIC, k_splits=ops.wgrad_splits(I, H, M))    # wgrad FFN up
    ops.gemm(dz, w1, dx, b_layout=1, epilogue=ops.EPI_ADD, aux=y)                            # dgrad FFN up + residual
    ops.gemm(dx, ctx, gwo, a_layout=1, b_layout=1, epilogue=ops.EPI_ATOMIC, k_splits=ops.wgrad_splits(H, H, M))  # wgrad out-proj
    ops.gemm(dx, wo, y, b_layout=1)                                                          # dgrad out-proj
    ops.attn_bwd(qkv, qkv, y, ctx, lse, dqkv, dqkv, ws, B, heads, S, S, q_col0=0, k_col0=H, v_col0=2 * H, dq_col0=0, dk_col0=H,
                 dv_col0=2 * H)                                                              # attention bwd
    ops.gemm(dqkv, x, gwqkv, a_layout=1, b_layout=1, epilogue=ops.EPI_ATOMIC, k_splits=ops.wgrad_splits(3 * H, H, M))  # wgrad QKV
    ops.gemm(dqkv, wqkv, dx, b_layout=1, epilogue=ops.EPI_ADD, aux=y)                        # dgrad QKV + residual
torch.cuda.synchronize()
print("done")
