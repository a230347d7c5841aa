"""A few launches of the fused halo conv at one shape (ncu target): python tools/conv_halo_one.py [B H C N]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coma_b200._lib import _stream, call  # noqa: E402

B, H, C, N = (int(v) for v in sys.argv[1:5]) if len(sys.argv) >= 5 else (4, 512, 128, 128)
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn((B, H, H, C), device=dev, generator=g).half()
wt = (torch.randn((N, 9 * C), device=dev, generator=g) * (9 * C) ** -0.5).half()
bias = torch.zeros(N, device=dev)
scale = torch.ones((B, C), device=dev)
shift = torch.zeros((B, C), device=dev)
out = torch.empty((B * H * H, N), dtype=torch.float16, device=dev)
stats = torch.empty((B * H * H // 32, N, 2), dtype=torch.float32, device=dev)
for _ in range(4):
    call("coma_conv3x3_halo_f16", x.data_ptr(), B, H, H, C, C, 0, scale.data_ptr(), shift.data_ptr(), 1, wt.data_ptr(), 9 * C, N, bias.data_ptr(), None, 0, None, 0,
         out.data_ptr(), N, stats.data_ptr(), _stream())
torch.cuda.synchronize()
print("done")
