"""G1 (tcgen05 GEMM) parity: fp16 operands, fp32 accumulation, checked against an fp32 torch matmul of the SAME fp16-rounded
operands (tolerance 1e-4 relative to the row scale, i.e. accumulation-order noise only; fp16 outputs add one rounding)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _ref(a, w, bias, res, act):
    y = a.float() @ w.float().t()
    if bias is not None:
        y = y + bias
    if res is not None:
        y = y + res.float()
    if act == 1:
        y = torch.nn.functional.silu(y)
    return y


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 128, 320), (4096, 320, 2880), (64, 1280, 1280), (200, 72, 88),
                                   (1000, 640, 5760), (128, 8, 2880), (77, 320, 768)])
def test_gemm_f32_out(dev, M, N, K):
    from coma_b200 import ops
    g = torch.Generator(device=dev).manual_seed(M + N + K)
    a = (torch.randn((M, K), device=dev, generator=g) * 0.5).half()
    w = (torch.randn((N, K), device=dev, generator=g) * K ** -0.5).half()
    out = ops.gemm_f16(a, w, out_dtype=torch.float32)
    ref = _ref(a, w, None, None, 0)
    scale = ref.abs().max().item()
    assert (out - ref).abs().max().item() <= 1e-4 * scale, (out - ref).abs().max().item() / scale


@pytest.mark.parametrize("act", [0, 1])
def test_gemm_epilogue(dev, act):
    from coma_b200 import ops
    g = torch.Generator(device=dev).manual_seed(act)
    M, N, K = 1024, 640, 960
    a = torch.randn((M, K), device=dev, generator=g).half()
    w = (torch.randn((N, K), device=dev, generator=g) * K ** -0.5).half()
    bias = torch.randn(N, device=dev, generator=g)
    res = torch.randn((M, N), device=dev, generator=g).half()
    ref = _ref(a, w, bias, res, act)
    out32 = ops.gemm_f16(a, w, bias, res, act, out_dtype=torch.float32)
    assert (out32 - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()
    out16 = ops.gemm_f16(a, w, bias, res, act, out_dtype=torch.float16)
    torch.testing.assert_close(out16.float(), ref, rtol=2e-3, atol=2e-3 * ref.abs().max().item() / 8)
    # strided A (a view into a wider buffer), as used for channel-concatenated activations
    wide = torch.randn((M, K + 64), device=dev, generator=g).half()
    out = ops.gemm_f16(wide[:, :K], w, out_dtype=torch.float32)
    assert (out - _ref(wide[:, :K], w, None, None, 0)).abs().max().item() <= 1e-4 * ref.abs().max().item()


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (300, 320, 320), (4096, 320, 320), (1000, 72, 88), (2048, 1280, 640), (77, 640, 768),
                                   (8192 + 40, 2560, 320), (512, 1288, 128)])
@pytest.mark.parametrize("res", [False, True])
def test_gemm_f16_out_tma_epilogue(dev, M, N, K, res):
    """fp16 outputs leave through the TMA epilogue (32x32 panels staged in swizzled shared memory, residual fetched by TMA):
    row / column tails are clipped by the tensor map, every panel of every tile must land exactly once."""
    from coma_b200.inpaint import nn
    g = torch.Generator(device=dev).manual_seed(M * 3 + N + K)
    a = (torch.randn((M, K), device=dev, generator=g) * 0.5).half()
    w = (torch.randn((N, K), device=dev, generator=g) * K ** -0.5).half()
    bias = torch.randn(N, device=dev, generator=g)
    r = torch.randn((M, N), device=dev, generator=g).half() if res else None
    rows_per_bias = 64
    brows = torch.randn(((M + rows_per_bias - 1) // rows_per_bias, N), device=dev, generator=g)
    out = torch.full((M + 2, N), 7.0, dtype=torch.float16, device=dev)      # guard rows: nothing may be written past M
    nn.gemm(a, w, bias, r, act=1, out=out[:M], bias_rows=brows, rows_per_bias=rows_per_bias, alpha=0.5)
    ref = 0.5 * (a.float() @ w.float().t()) + bias + brows.repeat_interleave(rows_per_bias, 0)[:M]
    if res:
        ref = ref + r.float()
    ref = torch.nn.functional.silu(ref)
    torch.testing.assert_close(out[:M].float(), ref, rtol=2e-3, atol=2e-3)
    assert (out[M:] == 7.0).all()


def test_gemm_f16_out_strided_and_batched(dev):
    """Output written into a column slice of a wider buffer (skip-concat target) and the batched form (nb1 x nb2)."""
    from coma_b200.inpaint import nn
    g = torch.Generator(device=dev).manual_seed(5)
    M, N, K = 640, 320, 192
    a = torch.randn((M, K), device=dev, generator=g).half()
    w = (torch.randn((N, K), device=dev, generator=g) * K ** -0.5).half()
    wide = torch.zeros((M, N + 64), dtype=torch.float16, device=dev)
    nn.gemm(a, w, out=wide[:, 64:])
    torch.testing.assert_close(wide[:, 64:].float(), a.float() @ w.float().t(), rtol=2e-3, atol=2e-3)
    assert (wide[:, :64] == 0).all()
    B, Hh, S, L, d = 2, 3, 200, 136, 64
    q = torch.randn((B, S, Hh * d), device=dev, generator=g).half()
    k = torch.randn((B, L, Hh * d), device=dev, generator=g).half()
    scores = torch.zeros((B, Hh, S, L), dtype=torch.float16, device=dev)
    nn.gemm_batched(q, Hh * d, d, S * Hh * d, k, Hh * d, d, L * Hh * d, scores, L, S * L, Hh * S * L, S, L, d, Hh, B, alpha=0.125)
    ref = 0.125 * torch.einsum("bshd,blhd->bhsl", q.float().view(B, S, Hh, d), k.float().view(B, L, Hh, d))
    torch.testing.assert_close(scores.float(), ref, rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize("M,N,K,out_dtype", [(512, 1280, 11520, torch.float16), (256, 640, 5760, torch.float32), (2048, 1280, 23040, torch.float16)])
def test_gemm_split_k(dev, M, N, K, out_dtype):
    """Deep-K problems with few output tiles run a deterministic split-K schedule (fp32 slabs + fixed-order finishing pass):
    same epilogue contract, bit-identical across repeats."""
    from coma_b200.inpaint import nn
    g = torch.Generator(device=dev).manual_seed(K)
    a = (torch.randn((M, K), device=dev, generator=g) * 0.5).half()
    w = (torch.randn((N, K), device=dev, generator=g) * K ** -0.5).half()
    bias = torch.randn(N, device=dev, generator=g)
    res = torch.randn((M, N), device=dev, generator=g).half()
    brows = torch.randn((M // 64, N), device=dev, generator=g)
    out = nn.gemm(a, w, bias, res, act=1, out_dtype=out_dtype, bias_rows=brows, rows_per_bias=64)
    ref = torch.nn.functional.silu(a.float() @ w.float().t() + bias + brows.repeat_interleave(64, 0) + res.float())
    tol = 2e-3 if out_dtype == torch.float16 else 1e-4
    assert (out.float() - ref).abs().max().item() <= tol * ref.abs().max().item()
    again = nn.gemm(a, w, bias, res, act=1, out_dtype=out_dtype, bias_rows=brows, rows_per_bias=64)
    assert torch.equal(out, again)


@pytest.mark.parametrize("M,C", [(1000, 320), (4096, 640), (300, 1280)])
def test_gemm_geglu_fused(dev, M, C):
    """Feed-forward projection with the GEGLU folded into the GEMM epilogue (interleaved weight rows) == projection (fp16) then
    value * gelu(gate), the diffusers GEGLU."""
    from coma_b200.inpaint import nn
    g = torch.Generator(device=dev).manual_seed(C)
    F = 4 * C
    x = torch.randn((M, C), device=dev, generator=g).half()
    w = (torch.randn((2 * F, C), device=dev, generator=g) * C ** -0.5).half()
    b = torch.randn(2 * F, device=dev, generator=g) * 0.1
    fused = nn.prep_geglu(w.cpu(), b.cpu(), dev)
    assert fused is not None
    out = nn.gemm_geglu(x, *fused)
    h = (x.float() @ w.float().t() + b).half().float()
    ref = h[:, :F] * torch.nn.functional.gelu(h[:, F:])
    torch.testing.assert_close(out.float(), ref, rtol=2e-3, atol=2e-3)
    unfused = nn.geglu(nn.gemm(x, nn.prep_linear(w, dev), b))
    torch.testing.assert_close(out.float(), unfused.float(), rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize("M,N,K,res", [(4096, 2560, 1280, True), (2048 + 256, 1280, 4096, False), (4096, 1024, 1024, True)])
def test_gemm_cta_pair_default_policy(dev, M, N, K, res):
    """160- / 256-wide tiles with an even number of M tiles and K >= 1024 run as CTA pairs (tcgen05.mma.cta_group::2, M = 256: each CTA stages
    its own A slab and half of the W slab, the leader issues for both): same epilogue contract as the single-CTA kernel."""
    from coma_b200._lib import call
    from coma_b200.inpaint import nn
    import ctypes
    bn, ks = ctypes.c_int(0), ctypes.c_int(0)
    call("coma_gemm_plan", M, N, K, 0, 0, ctypes.byref(bn), ctypes.byref(ks))
    assert bn.value in (160, 256) and ks.value == 1 and ((M + 127) // 128) % 2 == 0
    g = torch.Generator(device=dev).manual_seed(M + K)
    a = (torch.randn((M, K), device=dev, generator=g) * 0.5).half()
    w = (torch.randn((N, K), device=dev, generator=g) * K ** -0.5).half()
    bias = torch.randn(N, device=dev, generator=g)
    r = torch.randn((M, N), device=dev, generator=g).half() if res else None
    ref = a.float() @ w.float().t() + bias
    if res:
        ref = ref + r.float()
    out32 = nn.gemm(a, w, bias, r, out_dtype=torch.float32)
    assert (out32 - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()
    out16 = torch.full((M + 2, N), 7.0, dtype=torch.float16, device=dev)
    nn.gemm(a, w, bias, r, out=out16[:M])
    torch.testing.assert_close(out16[:M].float(), ref, rtol=2e-3, atol=2e-3)
    assert (out16[M:] == 7.0).all()


def test_gemm_cta_pair_forced(dev):
    """Every GEMM / implicit-conv parity case whose tile shape allows it, with COMA_GEMM_PAIR=1 (pairs regardless of K: short main loops,
    split-K slabs, GEGLU, residual / statistics epilogues, strided and multi-image conv tiles). The switch is read once per process."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, COMA_GEMM_PAIR="1")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_gpu_gemm.py"), os.path.join(here, "test_gpu_unet.py"), "-x", "-q", "-m", "gpu",
                        "-k", "(test_gemm and not cta_pair) or conv3x3_f16_out_tma or stride2 or from_conv_epilogue_stats or layernorm_folded"],
                       env=env, cwd=os.path.dirname(here), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0 and " passed" in r.stdout and "failed" not in r.stdout, r.stdout[-2000:]
