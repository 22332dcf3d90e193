"""Epilogue-bound or not?  GELU / RES32 / DGELU variants of the 2-CTA GEMM at the bench shapes under the measurement knobs."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spokennlp_b200 import lib, ops  # noqa: E402
from tools.gemm_sweep import timeit  # noqa: E402

M, H, I = 16384, 768, 3072
dev, f16 = "cuda", torch.float16
so = lib.load()
x = torch.randn(M, H, device=dev, dtype=f16)
w1 = torch.randn(I, H, device=dev, dtype=f16) * 0.02
w2 = torch.randn(H, I, device=dev, dtype=f16) * 0.02
b1, b2 = torch.zeros(I, device=dev), torch.zeros(H, device=dev)
h = torch.empty(M, I, device=dev, dtype=f16)
z = torch.empty(M, I, device=dev, dtype=f16)
pre = torch.empty(M, H, device=dev, dtype=torch.float32)
x32 = torch.randn(M, H, device=dev)
dx = torch.empty(M, H, device=dev, dtype=f16)
for dbg in (0, 3, 4, 7, 8, 15):
    so.b200_set_gemm_debug(dbg)
    t_plain = timeit(lambda: ops.gemm(x, w1, h))
    t_bias = timeit(lambda: ops.gemm(x, w1, h, epilogue=ops.EPI_BIAS, bias=b1))
    t_gelu = timeit(lambda: ops.gemm(x, w1, h, epilogue=ops.EPI_BIAS_GELU, bias=b1, out2=z))
    t_gelu1 = timeit(lambda: ops.gemm(x, w1, h, epilogue=ops.EPI_BIAS_GELU, bias=b1))
    t_res = timeit(lambda: ops.gemm(h, w2, pre, epilogue=ops.EPI_BIAS_RES32, bias=b2, aux=x32))
    t_st32 = timeit(lambda: ops.gemm(h, w2, pre))
    t_dg = timeit(lambda: ops.gemm(dx, w2, h, b_layout=1, epilogue=ops.EPI_DGELU, aux=z))
    t_add = timeit(lambda: ops.gemm(h, w1, dx, b_layout=1, epilogue=ops.EPI_ADD, aux=x))
    print(f"dbg={dbg:2d}: up plain {t_plain*1e6:6.1f} bias {t_bias*1e6:6.1f} gelu+d {t_gelu*1e6:6.1f} gelu {t_gelu1*1e6:6.1f} | down store32 {t_st32*1e6:6.1f} "
          f"res32 {t_res*1e6:6.1f} | dgelu(mul) {t_dg*1e6:6.1f} add {t_add*1e6:6.1f}  [us]", flush=True)
so.b200_set_gemm_debug(0)
