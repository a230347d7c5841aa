"""C2 halo-tile conv (+ fused GroupNorm affine / SiLU) against the implicit-GEMM conv + separate affine_act pass, on the UNet / VAE shapes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coma_b200._lib import _stream, call  # noqa: E402
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
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        gr.replay()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / n)
    return best


SHAPES = [(4, 512, 128, 128), (4, 512, 256, 128), (4, 256, 256, 256), (4, 256, 512, 256), (4, 128, 512, 512), (4, 64, 512, 512),
          (8, 64, 320, 320), (8, 64, 640, 320), (8, 32, 640, 640), (8, 32, 1280, 640)]
print("B  HxW   Cin->Cout |  implicit conv   affine pass |   halo (plain)   halo (fused)  | speed-up of the norm->SiLU->conv group")
for B, H, C, N in SHAPES:
    x = torch.randn((B, H, H, C), device=dev, generator=g).half()
    wt = (torch.randn((N, 9 * C), device=dev, generator=g) * (9 * C) ** -0.5).half()
    bias = torch.zeros(N, device=dev)
    scale = torch.ones((B, C), device=dev)
    shift = torch.zeros((B, C), device=dev)
    out = torch.empty((B * H * H, N), dtype=torch.float16, device=dev)
    y = torch.empty_like(x)
    stats = torch.empty((B * H * H // 32, N, 2), dtype=torch.float32, device=dev)
    sw = torch.zeros(1, dtype=torch.int32)

    def old_conv():
        call("coma_conv3x3_strided_f16", x.data_ptr(), B, H, H, C, C, 1, 1, wt.data_ptr(), 9 * C, N, bias.data_ptr(), None, 0, None, 0, out.data_ptr(), None, N,
             None, 0, stats.data_ptr(), None, _stream())

    def old_affine():
        call("coma_affine_act_f16", x.data_ptr(), B, H * H, C, C, scale.data_ptr(), shift.data_ptr(), 1, y.data_ptr(), C, _stream())

    def halo(fused):
        call("coma_conv3x3_halo_f16", x.data_ptr(), B, H, H, C, C, 0, scale.data_ptr() if fused else None, shift.data_ptr() if fused else None, 1, wt.data_ptr(),
             9 * C, N, bias.data_ptr(), None, 0, None, 0, out.data_ptr(), N, stats.data_ptr(), _stream())

    t_conv = t_aff = t_plain = t_fused = 1e9
    for _ in range(3):   # interleaved rounds, best of: clocks drift under the power cap
        t_conv, t_aff = min(t_conv, timeit(old_conv)), min(t_aff, timeit(old_affine))
        t_plain, t_fused = min(t_plain, timeit(lambda: halo(False))), min(t_fused, timeit(lambda: halo(True)))
    fl = 2.0 * B * H * H * 9 * C * N
    print(f"{B} {H:4d}^2 {C:4d}->{N:4d} | {t_conv * 1e3:7.1f} us {fl / t_conv / 1e9:6.0f} TF  {t_aff * 1e3:7.1f} us | {t_plain * 1e3:7.1f} us {fl / t_plain / 1e9:6.0f} TF"
          f"  {t_fused * 1e3:7.1f} us {fl / t_fused / 1e9:6.0f} TF | {(t_conv + t_aff) / t_fused:5.2f}x")
