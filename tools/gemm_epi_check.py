"""Correctness + timing of every gemm2 epilogue variant at the multi-tile bench shapes (M=16384), one at a time."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spokennlp_b200 import lib, ops  # noqa: E402
from tools.timing import timeit  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
M, H, I = int(os.environ.get("M", 16384)), 768, 3072
dev, f16 = "cuda", torch.float16
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s, sc=1.0: (torch.randn(*s, device=dev, generator=g) * sc)
x = rn(M, H).half()
w1 = rn(I, H, sc=0.05).half()
w2 = rn(H, I, sc=0.05).half()
b1, b2 = rn(I, sc=0.1), rn(H, sc=0.1)
hbig = rn(M, I).half()
x32 = rn(M, H)
aux_h = rn(M, H).half()
aux_i = rn(M, I).half()


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def check(name, fn, ref_fn, out):
    fn()
    torch.cuda.synchronize()
    r = rel(out, ref_fn())
    t = timeit(fn)
    print(f"{name:28s} rel={r:.2e}  {t * 1e6:7.1f} us", flush=True)


acc_up = x.float() @ w1.float().t()
acc_dn = hbig.float() @ w2.float().t()
o_i = torch.empty(M, I, device=dev, dtype=f16)
o_i2 = torch.empty(M, I, device=dev, dtype=f16)
o_h = torch.empty(M, H, device=dev, dtype=f16)
o_h32 = torch.empty(M, H, device=dev, dtype=torch.float32)
check("up STORE f16", lambda: ops.gemm(x, w1, o_i), lambda: acc_up, o_i)
check("up BIAS f16", lambda: ops.gemm(x, w1, o_i, epilogue=ops.EPI_BIAS, bias=b1), lambda: acc_up + b1, o_i)
check("up BIAS_GELU (no out2)", lambda: ops.gemm(x, w1, o_i, epilogue=ops.EPI_BIAS_GELU, bias=b1), lambda: torch.nn.functional.gelu(acc_up + b1), o_i)
check("up BIAS_GELU + derivative", lambda: ops.gemm(x, w1, o_i, epilogue=ops.EPI_BIAS_GELU, bias=b1, out2=o_i2),
      lambda: torch.nn.functional.gelu(acc_up + b1), o_i)
zf = (acc_up + b1).clone().requires_grad_(True)
torch.nn.functional.gelu(zf).sum().backward()
print(f"{'   saved derivative':28s} rel={rel(o_i2, zf.grad):.2e}", flush=True)
check("down STORE f32", lambda: ops.gemm(hbig, w2, o_h32), lambda: acc_dn, o_h32)
check("down BIAS_RES32", lambda: ops.gemm(hbig, w2, o_h32, epilogue=ops.EPI_BIAS_RES32, bias=b2, aux=x32), lambda: acc_dn + b2 + x32, o_h32)
check("down BIAS_RES f16", lambda: ops.gemm(hbig, w2, o_h, epilogue=ops.EPI_BIAS_RES, bias=b2, aux=aux_h), lambda: acc_dn + b2 + aux_h.float(), o_h)
w2t = w2  # [H, I] row-major == [K=H, N=I] for dX[M,I] = dY[M,H] @ W2[H,I]
acc_dg = x.float() @ w2t.float()
check("dgrad DGELU (mul aux)", lambda: ops.gemm(x, w2t, o_i, b_layout=1, epilogue=ops.EPI_DGELU, aux=aux_i), lambda: acc_dg * aux_i.float(), o_i)
acc_da = hbig.float() @ w1.float()     # dX[M,H] = dZ[M,I] @ W1[I,H]
check("dgrad ADD", lambda: ops.gemm(hbig, w1, o_h, b_layout=1, epilogue=ops.EPI_ADD, aux=aux_h), lambda: acc_da + aux_h.float(), o_h)
check("dgrad STORE", lambda: ops.gemm(hbig, w1, o_h, b_layout=1), lambda: acc_da, o_h)
print("done")
