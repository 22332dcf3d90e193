"""Where does the 2-CTA GEMM mainloop lose time?  Times the bench shapes with the measurement knobs of
b200_set_gemm_debug (1 skip A loads, 2 skip B loads, 4 skip MMA issue, 8 skip epilogue stores)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spokennlp_b200 import lib, ops  # noqa: E402


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def main():
    M = 16384
    dev, f16 = "cuda", torch.float16
    so = lib.load()
    shapes = [("ffn_up  N3072 K768 ", 3072, 768), ("ffn_down N768 K3072", 768, 3072), ("qkv     N2304 K768 ", 2304, 768)]
    for impl in (2, 1):
        ops.set_gemm_impl(impl)
        for name, N, K in shapes:
            a = torch.randn(M, K, device=dev, dtype=f16)
            w = torch.randn(N, K, device=dev, dtype=f16) * 0.02
            wt = w.t().contiguous()
            out = torch.empty(M, N, device=dev, dtype=f16)
            bias = torch.zeros(N, device=dev)
            for dbg in ((0, 1, 2, 3, 4, 7, 8, 11, 15) if impl == 2 else (0,)):
                so.b200_set_gemm_debug(dbg)
                t = timeit(lambda: ops.gemm(a, w, out))
                t2 = timeit(lambda: ops.gemm(a, wt, out, b_layout=1))
                t3 = timeit(lambda: ops.gemm(a, w, out, epilogue=ops.EPI_BIAS_GELU, bias=bias, out2=out)) if dbg in (0, 8) else float("nan")
                fl = 2.0 * M * N * K / 1e12
                print(f"impl{impl} {name} dbg={dbg:2d}: K-major {t * 1e6:7.1f} us {fl / t:7.0f} TF | B MN-major {t2 * 1e6:7.1f} us {fl / t2:7.0f} TF"
                      f" | gelu {t3 * 1e6:7.1f} us", flush=True)
            so.b200_set_gemm_debug(0)
        # wgrad shape
        dy = torch.randn(M, 3072, device=dev, dtype=f16)
        x = torch.randn(M, 768, device=dev, dtype=f16)
        gw = torch.zeros(3072, 768, device=dev)
        for dbg in ((0, 4, 8) if impl == 2 else (0,)):
            so.b200_set_gemm_debug(dbg)
            for sp in (3, 4, 5, 6, 8):
                t = timeit(lambda: ops.gemm(dy, x, gw, a_layout=1, b_layout=1, epilogue=ops.EPI_ATOMIC, k_splits=sp))
                print(f"impl{impl} wgrad 3072x768xT dbg={dbg} splits={sp}: {t * 1e6:7.1f} us {2.0 * M * 3072 * 768 / 1e12 / t:7.0f} TF", flush=True)
        so.b200_set_gemm_debug(0)


if __name__ == "__main__":
    main()
