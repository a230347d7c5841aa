"""Short-K projections of the UNet (K = 320 / 640) in isolation: time per launch from a CUDA graph of 10, interleaved rounds."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coma_b200.inpaint import nn  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
SHAPES = [(32768, 320, 320), (32768, 960, 320), (32768, 2560, 320), (8192, 640, 640), (8192, 1920, 640), (8192, 5120, 640), (2048, 1280, 1280)]
cases = []
for M, N, K in SHAPES:
    a = torch.randn((M, K), device=dev, generator=g).half()
    w = torch.randn((N, K), device=dev, generator=g).half()
    out = torch.empty((M, N), dtype=torch.float16, device=dev)
    res = torch.randn((M, N), device=dev, generator=g).half()
    bias = torch.randn(N, device=dev, generator=g)
    for _ in range(3):
        nn.gemm(a, w, bias, res, out=out)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(10):
            nn.gemm(a, w, bias, res, out=out)
    cases.append((M, N, K, gr, (a, w, out, res, bias)))
best = {}
for rnd in range(5):
    for M, N, K, gr, _ in cases:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gr.replay()
        e1.record()
        torch.cuda.synchronize()
        best[(M, N, K)] = min(best.get((M, N, K), 1e9), e0.elapsed_time(e1) / 10)
for (M, N, K), ms in best.items():
    print(f"{M:7d} {N:6d} {K:6d}: {ms * 1e3:8.1f} us  {2 * M * N * K / ms / 1e9:8.1f} TFLOP/s")
