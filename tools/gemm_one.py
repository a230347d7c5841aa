"""One GEMM shape with bias + residual, a few launches: the target of an ncu capture.  python tools/gemm_one.py M N K [res]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coma_b200.inpaint import nn  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
M, N, K = (int(v) for v in sys.argv[1:4])
use_res = len(sys.argv) > 4 and sys.argv[4] == "res"
a = torch.randn((M, K), device=dev, generator=g).half()
w = torch.randn((N, K), device=dev, generator=g).half()
out = torch.empty((M, N), dtype=torch.float16, device=dev)
res = torch.randn((M, N), device=dev, generator=g).half() if use_res else None
bias = torch.randn(N, device=dev, generator=g)
for _ in range(6):
    nn.gemm(a, w, bias, res, out=out)
torch.cuda.synchronize()
