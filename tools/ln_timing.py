"""LayerNorm forward / backward at the bench shape against a variant build:  B200_LIB=... python tools/ln_timing.py"""
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

M, H = 16384, 768
dev = "cuda"
pre = torch.randn(M, H, device=dev)
g, b = torch.ones(H, device=dev), torch.zeros(H, device=dev)
y = torch.empty(M, H, device=dev, dtype=torch.float16)
y32 = torch.empty(M, H, device=dev)
mean, rstd = torch.zeros(M, device=dev), torch.ones(M, device=dev)
t0 = timeit(lambda: ops.layernorm_fwd(pre, g, b, 1e-12, y=y, y32=y32, mean=mean, rstd=rstd))
dxl, dxd = torch.empty(M, H, device=dev, dtype=torch.float16), torch.empty(M, H, device=dev, dtype=torch.float16)
dgl, dbl, dbias = torch.zeros(H, device=dev), torch.zeros(H, device=dev), torch.zeros(H, device=dev)
drop = ops.Dropout(torch.tensor([7], dtype=torch.int32, device=dev), 5, 0.1)
t1 = timeit(lambda: ops.layernorm_bwd(y, pre, mean, rstd, g, dxl, dgl, dbl, dbias=dbias, dx_drop=dxd, drop=drop))
t2 = timeit(lambda: ops.layernorm_bwd(y, pre, mean, rstd, g, dxl, dgl, dbl, dbias=dbias))
print(json.dumps({"lib": os.path.basename(lib.LIB_PATH), "ln_fwd_us": t0 * 1e6, "ln_bwd_drop_us": t1 * 1e6, "ln_bwd_us": t2 * 1e6,
                  "ln_bwd_drop_GBps": M * H * 10 / t1 / 1e9, "ln_bwd_GBps": M * H * 8 / t2 / 1e9}))
