"""GEMM shapes of the encoder timed against a variant build of the library (operand-ring depth experiments).

    B200_LIB=tools/micro/libb200enc_s4.so python tools/gemm_stages.py        (on a B200)
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spokennlp_b200 import lib  # noqa: E402

if os.environ.get("B200_LIB"):
    lib.LIB_PATH = os.path.join(ROOT, os.environ["B200_LIB"])
    lib.is_stale = lambda: False
from spokennlp_b200 import ops  # noqa: E402
from tools.timing import timeit  # noqa: E402


def main():
    M, H, I = 16384, 768, 3072
    dev, f16 = "cuda", torch.float16
    x = torch.randn(M, H, device=dev, dtype=f16)
    w1 = torch.randn(I, H, device=dev, dtype=f16) * 0.02
    b1 = torch.zeros(I, device=dev)
    z = torch.empty(M, I, device=dev, dtype=f16)
    d = torch.empty(M, I, device=dev, dtype=f16)
    h = torch.randn(M, I, device=dev, dtype=f16)
    w2 = torch.randn(H, I, device=dev, dtype=f16) * 0.02
    b2 = torch.zeros(H, device=dev)
    acc = torch.randn(M, H, device=dev)
    dy = torch.randn(M, I, device=dev, dtype=f16)
    gw = torch.zeros(I, H, device=dev)
    row = {"lib": os.path.basename(lib.LIB_PATH)}
    fl = 2.0 * M * I * H / 1e12
    t = timeit(lambda: ops.gemm(x, w1, z, epilogue=ops.EPI_BIAS, bias=b1))
    row["ffn_up_bias_us"], row["ffn_up_bias_tflops"] = t * 1e6, fl / t
    t = timeit(lambda: ops.gemm(x, w1, z, epilogue=ops.EPI_BIAS_GELU, bias=b1, out2=d))
    row["ffn_up_gelu_us"], row["ffn_up_gelu_tflops"] = t * 1e6, fl / t
    t = timeit(lambda: ops.gemm_resadd(h, w2, acc, b2))
    row["ffn_down_resadd_us"], row["ffn_down_resadd_tflops"] = t * 1e6, fl / t
    t = timeit(lambda: ops.gemm(dy, x, gw, a_layout=1, b_layout=1, epilogue=ops.EPI_ATOMIC, k_splits=ops.wgrad_splits(I, H, M)))
    row["wgrad_ffn_up_us"], row["wgrad_ffn_up_tflops"] = t * 1e6, fl / t
    print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
