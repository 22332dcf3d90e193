"""The 2-CTA GEMM kernel on square and long-K shapes against torch.matmul (cuBLAS) on the same operands: separates a mainloop
inefficiency from shape effects (short K = 12 k-blocks per tile on the encoder's shapes).

    python tools/gemm_big.py        (on a B200)
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spokennlp_b200 import ops  # noqa: E402
from tools.timing import timeit  # noqa: E402


def main():
    dev, f16 = "cuda", torch.float16
    for M, N, K in ((8192, 8192, 8192), (16384, 3072, 768), (16384, 3072, 3072), (16384, 3072, 12288), (16384, 768, 3072), (16384, 2304, 768)):
        a = torch.randn(M, K, device=dev, dtype=f16)
        b = torch.randn(N, K, device=dev, dtype=f16) * 0.02
        out = torch.empty(M, N, device=dev, dtype=f16)
        fl = 2.0 * M * N * K / 1e12
        t0 = timeit(lambda: ops.gemm(a, b, out))
        t1 = timeit(lambda: torch.matmul(a, b.t(), out=out))
        ab, bb = a.to(torch.bfloat16), b.to(torch.bfloat16)
        ob = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        t2 = timeit(lambda: torch.matmul(ab, bb.t(), out=ob))
        print(json.dumps({"shape": [M, N, K], "ours_us": t0 * 1e6, "ours_tflops": fl / t0, "cublas_f16_tflops": fl / t1, "cublas_bf16_tflops": fl / t2}), flush=True)


if __name__ == "__main__":
    main()
