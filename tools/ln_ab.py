"""A/B of the two LayerNorm-backward kernels at the bench shape ([16384, 768], fp32 pre-LN input), with the variants the
training step uses (second upstream gradient, dropout copy).  b200_set_gemm_debug bit 0x200000 selects the first one."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spokennlp_b200 import lib, ops  # noqa: E402
from tools.gemm_sweep import timeit  # noqa: E402

so = lib.load()
torch.manual_seed(0)
for M, H in ((16384, 768), (1000, 128), (4099, 1024)):
    dy = (torch.randn(M, H, device="cuda") * 0.1).half()
    dy2 = (torch.randn(M, H, device="cuda") * 0.1).half()
    x = torch.randn(M, H, device="cuda") * 2 + 0.3
    mean = x.mean(1).contiguous()
    rstd = (x.var(1, unbiased=False) + 1e-12).rsqrt().contiguous()
    gamma = torch.randn(H, device="cuda")
    seed = torch.tensor([77], dtype=torch.int32, device="cuda")
    for use_dy2, use_drop in ((False, False), (True, False), (True, True)):
        res = {}
        for v1 in (True, False):
            so.b200_set_gemm_debug(0x200000 if v1 else 0)
            dx = torch.zeros(M, H, device="cuda", dtype=torch.float16)
            dxd = torch.zeros(M, H, device="cuda", dtype=torch.float16) if use_drop else None
            dg, db, dbias = (torch.zeros(H, device="cuda") for _ in range(3))
            kw = dict(dy2=dy2 if use_dy2 else None, dbias=dbias, dx_drop=dxd, drop=ops.Dropout(seed, 5, 0.1) if use_drop else None)
            ops.layernorm_bwd(dy, x, mean, rstd, gamma, dx, dg, db, **kw)
            torch.cuda.synchronize()
            out = [t.clone() for t in (dx, dg, db, dbias)] + ([dxd.clone()] if use_drop else [])
            t = timeit(lambda: ops.layernorm_bwd(dy, x, mean, rstd, gamma, dx, dg, db, **kw))
            res[v1] = (out, t)
        so.b200_set_gemm_debug(0)
        rel = [float((a.float() - b.float()).norm() / (b.float().norm() + 1e-30)) for a, b in zip(res[False][0], res[True][0])]
        byt = M * H * (4 + 2 + 2 + (2 if use_dy2 else 0) + (2 if use_drop else 0))
        print(f"[{M},{H}] dy2={use_dy2} drop={use_drop}: v1 {res[True][1] * 1e6:6.1f} us  v2 {res[False][1] * 1e6:6.1f} us "
              f"({byt / res[False][1] / 1e9:5.0f} GB/s) | v2 vs v1 rel: " + " ".join(f"{r:.1e}" for r in rel), flush=True)
