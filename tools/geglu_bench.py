"""Fused GEGLU projection (GEMM + value * gelu(gate) epilogue) on the UNet's three feed-forward shapes: time and TFLOP/s."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coma_b200.inpaint import nn  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(n):
            fn()
    gr.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        gr.replay()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / n)
    return best


for M, C in ((32768, 320), (8192, 640), (2048, 1280)):
    x = torch.randn((M, C), device=dev, generator=g).half()
    w = torch.randn((8 * C, C), generator=torch.Generator().manual_seed(1)) * C ** -0.5
    b = torch.zeros(8 * C)
    wp, bp = nn.prep_geglu(w, b, dev)
    t = timeit(lambda: nn.gemm_geglu(x, wp, bp))
    t2 = timeit(lambda: nn.gemm(x, wp, bp))
    print(f"M={M:6d} C={C:5d} -> {8 * C:6d}: GEGLU {t * 1e3:7.1f} us {2.0 * M * C * 8 * C / t / 1e9:6.0f} TFLOP/s | plain GEMM, same weights {t2 * 1e3:7.1f} us")
