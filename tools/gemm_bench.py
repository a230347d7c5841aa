"""TFLOP/s of the tcgen05 GEMM / implicit conv on the shapes the UNet and VAE actually run (CUDA events, 1 GPU)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coma_b200.inpaint import nn  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)


def timeit(fn, n=10):
    """n back-to-back launches replayed from a CUDA graph (no host launch gaps), best of 3 replays."""
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
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        gr.replay()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / n)
    return best


print("--- plain GEMM (M, N, K)")
SHAPES = [(32768, 320, 320), (32768, 2560, 320), (32768, 320, 1280), (8192, 640, 640), (8192, 5120, 640), (8192, 640, 2560),
                (2048, 1280, 1280), (2048, 10240, 1280), (2048, 1280, 5120), (32768, 320, 960), (512, 1280, 1280), (616, 1280, 768),
          (8192, 8192, 8192)]
if len(sys.argv) > 1 and sys.argv[1] == "one":       # a single shape, a few launches: the target of an ncu capture
    M, N, K = (int(v) for v in sys.argv[2:5])
    a = torch.randn((M, K), device=dev, generator=g).half()
    w = torch.randn((N, K), device=dev, generator=g).half()
    out = torch.empty((M, N), dtype=torch.float16, device=dev)
    for _ in range(6):
        nn.gemm(a, w, out=out)
    torch.cuda.synchronize()
    sys.exit(0)
for M, N, K in SHAPES:
    a = torch.randn((M, K), device=dev, generator=g).half()
    w = torch.randn((N, K), device=dev, generator=g).half()
    out = torch.empty((M, N), dtype=torch.float16, device=dev)
    ms = timeit(lambda: nn.gemm(a, w, out=out))
    print(f"{M:7d} {N:6d} {K:6d}: {ms * 1e3:9.1f} us  {2 * M * N * K / ms / 1e9:8.1f} TFLOP/s")
print("--- implicit conv3x3 (B, H, C -> Cout)")
for B, H, C, Co in [(8, 64, 320, 320), (8, 64, 640, 320), (8, 64, 960, 320), (8, 32, 640, 640), (8, 32, 1280, 640), (8, 16, 1280, 1280),
                    (8, 16, 2560, 1280), (8, 8, 1280, 1280), (4, 512, 128, 128), (4, 256, 256, 256), (4, 256, 512, 512), (4, 128, 512, 512),
                    (4, 64, 512, 512)]:
    x = nn.new_act(B, H, H, C, dev)
    x.t.copy_(torch.randn((B * H * H, C), device=dev, generator=g).half())
    w = (torch.randn((Co, 9 * C), device=dev, generator=g) * 0.01).half()
    ms = timeit(lambda: nn.conv3x3(x, w, None), n=5)
    print(f"B{B} {H:4d}^2 {C:5d}->{Co:5d}: {ms * 1e3:9.1f} us  {2 * B * H * H * Co * 9 * C / ms / 1e9:8.1f} TFLOP/s")
