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
