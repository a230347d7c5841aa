"""Layer-level building blocks of the inpainting UNet / VAE on the B200 kernels (include/coma_b200.h, G1 + U*).

Activations are NHWC fp16: an `Act` wraps a 2-D tensor [B*H*W, C] (row stride may exceed C) plus its (B, H, W).
torch is used for allocation and views only; every arithmetic op is a coma_b200 kernel.
"""
import ctypes
import os
from dataclasses import dataclass

import torch

from .. import _lib
from .._lib import GemmArgs, _ptr, _stream, call

F16, F32 = torch.float16, torch.float32


def rup(x, m=8):
    return (x + m - 1) // m * m


@dataclass
class Act:
    t: torch.Tensor  # [B*H*W, C] fp16, stride(1) == 1
    B: int
    H: int
    W: int
    stats: torch.Tensor = None   # [B*H*W/stats_rows, C, 2] f32 partial (sum, sumsq) left by the producing conv's epilogue, or None
    stats_rows: int = 32         # pixels per partial-sum row (32: implicit-GEMM conv, 128: halo-tile conv)

    @property
    def C(self):
        return self.t.shape[1]

    @property
    def ld(self):
        return self.t.stride(0)

    @property
    def M(self):
        return self.t.shape[0]


def new_act(B, H, W, C, device, dtype=F16):
    """Fresh activation; channel counts that are not multiples of 8 get zero-padded rows (TMA needs 16-byte row strides)."""
    ld = rup(C)
    buf = torch.zeros((B * H * W, ld), dtype=dtype, device=device) if ld != C else torch.empty((B * H * W, C), dtype=dtype, device=device)
    return Act(buf[:, :C], B, H, W)


SPLITK_WS_ELEMS = 8 << 20   # 32 MB of fp32 partial slabs per device: enough for the 8x8 / 16x16 levels, small enough for L2
_SPLITK_WS = {}


def splitk_workspace(device):
    """Per-device fp32 scratch for the deterministic split-K schedule (deep-K problems with too few tiles for 148 SMs).
    Allocated once, outside any CUDA-graph capture (the eager warm-up of `Graphed` runs first); launches on one stream
    serialise on it."""
    key = (torch.device(device).index, torch.cuda.current_stream(device).cuda_stream)   # concurrent streams must not share slabs
    ws = _SPLITK_WS.get(key)
    if ws is None:
        ws = _SPLITK_WS[key] = torch.empty(SPLITK_WS_ELEMS, dtype=F32, device=device)
    return ws


def _set_ln(g, ln, M, N, K):
    """ln = (stats, c1): stats [M, 2] f32 = layernorm_stats(a), or [M, K/32, 2] f32 = the per-panel sums the producer of `a` left (ln_out)."""
    st, c1 = ln
    assert st.dtype == F32 and st.is_contiguous() and c1.dtype == F32 and c1.numel() == N and st.shape[0] == M
    if st.dim() == 3:
        assert st.shape[1] * 32 == K and st.shape[2] == 2
        g.ln_partials_in, g.ln_eps = st.data_ptr(), 1e-5
    else:
        assert st.shape == (M, 2)
        g.ln_row_stats = st.data_ptr()
    g.ln_c1 = c1.data_ptr()


def ln_partials(M, N, device):
    """Buffer for gemm(..., ln_out=): (sum, sumsq) of every 32-column panel of every output row."""
    assert N % 32 == 0
    return torch.empty((M, N // 32, 2), dtype=F32, device=device)


def gemm(a, w, bias=None, residual=None, act=0, out=None, out_dtype=F16, bias_rows=None, rows_per_bias=0, alpha=1.0, K=None, ln=None, ln_out=None):
    """out[M,N] = act(alpha * a @ w[:, :K].T + bias + bias_rows[m // rows_per_bias] + residual). a [M,K] f16, w [N,>=K] f16.
    ln = (stats, c1): a LayerNorm of `a` folded into w / bias (fold_layernorm; stats: see _set_ln). ln_out = ln_partials(M, N) buffer: the
    epilogue leaves the LayerNorm partial sums of the OUTPUT rows for the next folded projection (no statistics kernel at all)."""
    M = a.shape[0]
    K = a.shape[1] if K is None else K
    N = w.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    g = GemmArgs()
    g.A, g.lda, g.W, g.ldw = a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0)
    g.M, g.N, g.K, g.nb1, g.nb2 = M, N, K, 1, 1
    g.out_f16 = out.data_ptr() if out.dtype == F16 else None
    g.out_f32 = out.data_ptr() if out.dtype == F32 else None
    g.ldo = out.stride(0)
    if residual is not None:
        assert residual.dtype == F16 and residual.stride(0) == out.stride(0) and residual.shape == out.shape
        g.residual = residual.data_ptr()
    g.bias = _ptr(bias)
    if bias_rows is not None:
        assert bias_rows.dtype == F32 and bias_rows.stride(1) == 1 and bias_rows.shape[1] == N and rows_per_bias > 0
        g.bias_rows, g.rows_per_bias, g.bias_rows_ld = bias_rows.data_ptr(), rows_per_bias, bias_rows.stride(0)
    g.alpha, g.act = alpha, act
    if ln is not None:
        _set_ln(g, ln, M, N, K)
    if ln_out is not None:
        assert ln_out.shape == (M, N // 32, 2) and ln_out.dtype == F32 and ln_out.is_contiguous() and out.dtype == F16
        g.ln_partials_out = ln_out.data_ptr()
    ws = splitk_workspace(a.device)
    g.workspace, g.workspace_elems = ws.data_ptr(), ws.numel()
    with torch.cuda.device(a.device):
        call("coma_gemm_f16_ex", ctypes.addressof(g), _stream())
    return out


def gemm_batched(A, lda, a_s1, a_s2, W, ldw, w_s1, w_s2, out, ldo, o_s1, o_s2, M, N, K, nb1, nb2, alpha=1.0):
    g = GemmArgs()
    g.A, g.lda, g.a_s1, g.a_s2 = A.data_ptr(), lda, a_s1, a_s2
    g.W, g.ldw, g.w_s1, g.w_s2 = W.data_ptr(), ldw, w_s1, w_s2
    g.out_f16 = out.data_ptr() if out.dtype == F16 else None
    g.out_f32 = out.data_ptr() if out.dtype == F32 else None
    g.ldo, g.o_s1, g.o_s2 = ldo, o_s1, o_s2
    g.M, g.N, g.K, g.nb1, g.nb2, g.alpha, g.act = M, N, K, nb1, nb2, alpha, 0
    with torch.cuda.device(A.device):
        call("coma_gemm_f16_ex", ctypes.addressof(g), _stream())
    return out


SMALL_N_CONV = True     # conv_out layers (Cout <= 4) on the direct fused kernel C1 (A/B switch for tests and tuning)
HALO_CONV = os.environ.get("COMA_NO_HALO_CONV") is None   # norm -> SiLU -> conv on the fused halo-tile kernel C2 (A/B switch)
_GN_COUNTERS = {}
FUSED_GN_STATS = True   # convolutions leave GroupNorm partial sums for their consumer (A/B switch for tests and tuning)


def gn_affine(x: Act, gamma, beta, groups, eps):
    """GroupNorm statistics folded with the affine: gn(x) = x*scale[b,c] + shift[b,c] (fp32 [B,C] each). One launch — and no
    pass over x at all when the convolution that produced x left its per-block partial sums (`x.stats`)."""
    dev = x.t.device
    if FUSED_GN_STATS and x.stats is not None:
        scale = torch.empty((x.B, x.C), dtype=F32, device=dev)
        shift = torch.empty((x.B, x.C), dtype=F32, device=dev)
        with torch.cuda.device(dev):
            call("coma_groupnorm_from_stats_rb_f32", x.stats.data_ptr(), x.B, x.H * x.W, x.stats_rows, x.C, groups, float(eps), _ptr(gamma), _ptr(beta),
                 None, None, scale.data_ptr(), shift.data_ptr(), _stream())
        return scale, shift
    ws = torch.empty(2 * groups * (4 * 148 + x.B), dtype=torch.float64, device=dev)
    ckey = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    cnt = _GN_COUNTERS.get(ckey)
    if cnt is None or cnt.numel() < x.B:   # ticket counters: zero once, every launch leaves them zero
        cnt = _GN_COUNTERS[ckey] = torch.zeros(max(1024, x.B), dtype=torch.int32, device=dev)
    scale = torch.empty((x.B, x.C), dtype=F32, device=dev)
    shift = torch.empty((x.B, x.C), dtype=F32, device=dev)
    with torch.cuda.device(dev):
        call("coma_groupnorm_affine_f16", x.t.data_ptr(), x.B, x.H * x.W, x.C, x.ld, groups, float(eps), _ptr(gamma), _ptr(beta),
             ws.data_ptr(), cnt.data_ptr(), None, None, scale.data_ptr(), shift.data_ptr(), _stream())
    return scale, shift


def affine_act(x: Act, scale, shift, act=0):
    y = new_act(x.B, x.H, x.W, x.C, x.t.device)
    with torch.cuda.device(x.t.device):
        call("coma_affine_act_f16", x.t.data_ptr(), x.B, x.H * x.W, x.C, x.ld, scale.data_ptr(), shift.data_ptr(), act,
             y.t.data_ptr(), y.ld, _stream())
    return y


def conv_out_hw(H, W, stride, pad, up):
    Hin, Win = (2 * H, 2 * W) if up else (H, W)
    if stride == 1:
        return Hin, Win
    return ((Hin + 2 - 3) // 2 + 1, (Win + 2 - 3) // 2 + 1) if pad else ((Hin + 1 - 3) // 2 + 1, (Win + 1 - 3) // 2 + 1)


def _tiles_128(H, W):
    tw = min(W, 128)
    th = min(128 // tw, H)
    return tw * th * (128 // (tw * th)) == 128 and W % tw == 0 and H % th == 0


def upsample2x_affine_act(x: Act, scale, shift, act):
    y = new_act(x.B, 2 * x.H, 2 * x.W, x.C, x.t.device)
    with torch.cuda.device(x.t.device):
        call("coma_upsample2x_affine_act_f16", x.t.data_ptr(), x.B, x.H, x.W, x.C, x.ld, _ptr(scale), _ptr(shift), act,
             y.t.data_ptr(), y.ld, _stream())
    return y


IMPLICIT_CONV = True


def conv3x3(x: Act, w, bias, stride=1, pad=1, up=False, gn=None, act=0, residual=None, bias_rows=None, out_dtype=F16, stats=False):
    """3x3 convolution on the tensor cores. w: [Cout, ld >= 9*Cin] f16 with K order (ky, kx, cin).
    stride 1, Cin % 64 == 0: implicit GEMM (shifted TMA tiles, no im2col matrix) on the pre-activated tensor;
    otherwise (stride 2, tiny Cin): im2col with the GroupNorm affine + SiLU applied while gathering, then GEMM."""
    Ho, Wo = conv_out_hw(x.H, x.W, stride, pad, up)
    if SMALL_N_CONV and w.shape[0] <= 4 and stride == 1 and pad == 1 and not up and x.C % 8 == 0 and residual is None and bias_rows is None:
        # C1: a handful of output channels (VAE / UNet conv_out): direct halo-tiled kernel with the GroupNorm affine + SiLU of the
        # input fused in — no normalised intermediate tensor, no 64-wide tensor-core tile for 3 columns
        scale, shift = gn if gn is not None else (None, None)
        out = new_act(x.B, Ho, Wo, w.shape[0], x.t.device, out_dtype)
        with torch.cuda.device(x.t.device):
            call("coma_conv3x3_small_n_f16", x.t.data_ptr(), x.B, x.H, x.W, x.C, x.ld, _ptr(scale), _ptr(shift), act if gn is not None else 0,
                 w.data_ptr(), w.stride(0), w.shape[0], _ptr(bias), out.t.data_ptr() if out_dtype == F32 else None,
                 out.t.data_ptr() if out_dtype == F16 else None, out.t.stride(0), _stream())
        return out
    if (HALO_CONV and (gn is not None or up) and stride == 1 and pad == 1 and out_dtype == F16 and Ho % 16 == 0 and Wo % 8 == 0
            and x.C % 64 == 0 and w.shape[0] % 64 == 0 and Ho * Wo >= 1024 and x.B * (Ho // 16) * (Wo // 8) * max(1, w.shape[0] // 256) >= 128):
        # C2: norm -> SiLU -> conv (and nearest x2 upsample -> conv) in one kernel (halo tiles, the affine + activation / the upsampling
        # applied while the tile is staged): no normalised or upsampled tensor in HBM. Faster than affine_act + implicit GEMM on every level with >= 128 pixel tiles (tools/conv_halo_bench.py:
        # 1.06-1.26x per layer); the 16^2 / 8^2 levels (a handful of tiles, split-K) stay on the implicit-GEMM path.
        N = w.shape[0]
        out = new_act(x.B, Ho, Wo, N, x.t.device, F16)
        if residual is not None:
            assert residual.dtype == F16 and residual.stride(0) == out.t.stride(0)
        want = stats and FUSED_GN_STATS
        st = torch.empty((x.B * Ho * Wo // 128, N, 2), dtype=F32, device=x.t.device) if want else None   # one row per 16 x 8 pixel tile
        with torch.cuda.device(x.t.device):
            call("coma_conv3x3_halo_f16", x.t.data_ptr(), x.B, Ho, Wo, x.C, x.ld, int(up), None if gn is None else _ptr(gn[0]),
                 None if gn is None else _ptr(gn[1]), act if gn is not None else 0, w.data_ptr(), w.stride(0), N,
                 _ptr(bias), None if bias_rows is None else bias_rows.data_ptr(), 0 if bias_rows is None else bias_rows.stride(0),
                 None if residual is None else residual.data_ptr(), 0, out.t.data_ptr(), out.t.stride(0), None if st is None else st.data_ptr(),
                 _stream())
        if st is not None:
            out.stats, out.stats_rows = st, 128
        return out
    strided_ok = stride == 2 and not up and gn is None and Ho * Wo >= 128   # element-strided TMA tiles
    if IMPLICIT_CONV and ((stride == 1 and pad == 1) or strided_ok) and x.C % 64 == 0 and _tiles_128(Ho, Wo):
        scale, shift = gn if gn is not None else (None, None)
        if up:
            xa = upsample2x_affine_act(x, scale, shift, act if gn is not None else 0)
        elif gn is not None:
            xa = affine_act(x, scale, shift, act)
        else:
            xa = x
        N = w.shape[0]
        out = new_act(x.B, Ho, Wo, N, x.t.device, out_dtype)
        if residual is not None:
            assert residual.dtype == F16 and residual.stride(0) == out.t.stride(0)
        ws = splitk_workspace(x.t.device)
        want = stats and FUSED_GN_STATS and out_dtype == F16 and (x.B * Ho * Wo) % 32 == 0 and (Ho * Wo) % 32 == 0
        st = torch.empty((x.B * Ho * Wo // 32, N, 2), dtype=F32, device=x.t.device) if want else None
        written = ctypes.c_int(0)
        with torch.cuda.device(x.t.device):
            call("coma_conv3x3_strided_f16", xa.t.data_ptr(), xa.B, xa.H, xa.W, xa.C, xa.ld, stride, 1 if stride == 1 else int(bool(pad)),
                 w.data_ptr(), w.stride(0), N, _ptr(bias),
                 None if bias_rows is None else bias_rows.data_ptr(), 0 if bias_rows is None else bias_rows.stride(0),
                 None if residual is None else residual.data_ptr(), 0,
                 out.t.data_ptr() if out_dtype == F16 else None, out.t.data_ptr() if out_dtype == F32 else None, out.t.stride(0),
                 ws.data_ptr(), ws.numel(), None if st is None else st.data_ptr(), ctypes.byref(written), _stream())
        if st is not None and written.value:
            out.stats = st
        return out
    K = 9 * x.C
    cols = torch.empty((x.B * Ho * Wo, rup(K)), dtype=F16, device=x.t.device)
    scale, shift = gn if gn is not None else (None, None)
    with torch.cuda.device(x.t.device):
        call("coma_im2col3x3_f16", x.t.data_ptr(), x.B, x.H, x.W, x.C, x.ld, stride, pad, int(up), _ptr(scale), _ptr(shift),
             act if gn is not None else 0, cols.data_ptr(), cols.stride(0), _stream())
    N = w.shape[0]
    out = new_act(x.B, Ho, Wo, N, x.t.device, out_dtype)
    gemm(cols, w, bias, None if residual is None else residual, 0, out.t, bias_rows=bias_rows, rows_per_bias=Ho * Wo, K=rup(K))
    return out


def layernorm(x2d, gamma, beta, eps=1e-5):
    y = torch.empty((x2d.shape[0], x2d.shape[1]), dtype=F16, device=x2d.device)
    with torch.cuda.device(x2d.device):
        call("coma_layernorm_f16", x2d.data_ptr(), x2d.shape[0], x2d.shape[1], x2d.stride(0), gamma.data_ptr(), beta.data_ptr(),
             float(eps), y.data_ptr(), y.stride(0), _stream())
    return y


# LayerNorm folded into the projection that consumes it (gamma into the weights, beta into the bias, (rstd, -rstd * mean) applied per row in
# the GEMM epilogue; statistics from layernorm_stats or from the producer GEMM's epilogue). Built, parity-tested — and OFF by default:
# measured on the SD-1.5 UNet (B200, tools/unet_breakdown.py) the 48 LayerNorm launches (0.63 ms) disappear but the short-K projections are
# epilogue-bound, and the extra epilogue work costs as much: 9.92 ms unfused, 9.79 ms with the statistics kernel, 9.98 ms with producer-side
# statistics. COMA_FUSED_LN=1 turns it on.
FUSED_LN = os.environ.get("COMA_FUSED_LN") == "1"


def layernorm_stats(x2d, eps=1e-5):
    """[M, 2] f32 = (rstd, -rstd * mean) per row: the per-row inputs of a LayerNorm folded into the consuming GEMM (no normalised tensor)."""
    st = torch.empty((x2d.shape[0], 2), dtype=F32, device=x2d.device)
    with torch.cuda.device(x2d.device):
        call("coma_layernorm_stats_f16", x2d.data_ptr(), x2d.shape[0], x2d.shape[1], x2d.stride(0), float(eps), st.data_ptr(), _stream())
    return st


def fold_layernorm(w, b, gamma, beta):
    """LayerNorm(x) @ w.T + b  ==  rstd * (x @ w'.T) - rstd * mean * c1 + b'  with w' = w * gamma, c1 = row sums of w' (as stored: fp16),
    b' = w @ beta + b. fp32 host tensors in -> (w' fp32 [N,K], b' fp32 [N]); c1 is taken from the prepared fp16 weight (ln_c1)."""
    w, gamma, beta = w.float().reshape(w.shape[0], -1), gamma.float(), beta.float()
    b2 = w @ beta + (b.float() if b is not None else 0.0)
    return w * gamma[None, :], b2


def ln_c1(w_prepared):
    """Column-sum vector of a prepared (fp16, zero-padded) folded weight: c1[n] = sum_k w'[n, k] in fp32."""
    return w_prepared.float().sum(1).contiguous()


def prep_geglu(w, b, device):
    """Feed-forward projection [2F, C] (rows: F values, then F gates) -> rows interleaved in blocks of 32 (32 values, their 32
    gates, ...) so one 64-column accumulator group of the GEMM holds value and gate of the same features (fused GEGLU
    epilogue). Returns (weight [2F, rup(C)] f16, bias [2F] f32) or None when the fused kernel does not apply."""
    F = w.shape[0] // 2
    if (2 * F) % 256 != 0:
        return None
    blocks = torch.arange(F // 32).view(-1, 1, 1) * 32 + torch.arange(32).view(1, 1, 32)   # feature index per slot
    src = torch.cat([blocks, blocks + F], dim=1).reshape(-1)       # new row r <- old row src[r]
    return prep_linear(w[src], device), prep_vec(b[src], device)


def gemm_geglu(a, w_il, b_il, ln=None):
    """out[M, F] = value * gelu(gate) of the interleaved projection (see prep_geglu), one kernel. ln: as in gemm."""
    M, K = a.shape
    N = w_il.shape[0]
    out = torch.empty((M, N // 2), dtype=F16, device=a.device)
    g = GemmArgs()
    g.A, g.lda, g.W, g.ldw = a.data_ptr(), a.stride(0), w_il.data_ptr(), w_il.stride(0)
    g.M, g.N, g.K, g.nb1, g.nb2 = M, N, K, 1, 1
    g.out_f16, g.ldo, g.bias, g.alpha, g.geglu = out.data_ptr(), out.stride(0), _ptr(b_il), 1.0, 1
    if ln is not None:
        _set_ln(g, ln, M, N, K)
    with torch.cuda.device(a.device):
        call("coma_gemm_f16_ex", ctypes.addressof(g), _stream())
    return out


def geglu(h):
    C = h.shape[1] // 2
    y = torch.empty((h.shape[0], C), dtype=F16, device=h.device)
    with torch.cuda.device(h.device):
        call("coma_geglu_f16", h.data_ptr(), h.shape[0], C, h.stride(0), y.data_ptr(), y.stride(0), _stream())
    return y


def silu(x):
    y = torch.empty_like(x)
    with torch.cuda.device(x.device):
        call("coma_silu_f16", x.data_ptr(), x.numel(), y.data_ptr(), _stream())
    return y


def timestep_embedding(t, dim):
    out = torch.empty((t.shape[0], dim), dtype=F16, device=t.device)
    with torch.cuda.device(t.device):
        call("coma_timestep_embedding_f16", t.data_ptr(), t.shape[0], dim, out.data_ptr(), _stream())
    return out


FUSED_ATTENTION = True


def project_kv(xkv, B, L, wkv, heads):
    """K and V^T of an attention layer from ONE projection GEMM over the concatenated [2C, Ckv] weight: returns
    (k [B*L, C] view with row stride 2C, vt [B, heads, d, Lp]). The UNet's cross-attention context is constant over the
    denoising loop, so the pipeline computes these once per call instead of once per step."""
    C = wkv.shape[0] // 2
    d = C // heads
    kv = gemm(xkv, wkv)
    Lp = rup(L)
    vt = torch.empty((B, heads, d, Lp), dtype=F16, device=xkv.device)
    v = kv[:, C:]
    with torch.cuda.device(xkv.device):
        call("coma_transpose_heads_f16", v.data_ptr(), B, L, heads, d, v.stride(0), vt.data_ptr(), Lp, _stream())
    return kv[:, :C], vt


def attention(xq, xkv, B, S, L, wq, wk, wv, wo, bo, heads, residual, wqkv=None, wkv=None, kv=None, ln=None, ln_out=None):
    """softmax(Q K^T / sqrt(d)) V followed by the output projection (+bias +residual). xq [B*S, C], xkv [B*L, Ckv] f16.
    wqkv ([3C, C], self-attention) / wkv ([2C, Ckv]) are the row-concatenated projection weights: one GEMM instead of three /
    two; kv = (k, vt) supplies precomputed keys / transposed values (see project_kv). ln = (row_stats, c1, bias'): xq is the
    UN-normalised input and the LayerNorm is folded into wqkv (self-attention) / wq (fold_layernorm); ln_out: see gemm (output projection)."""
    lnk = dict(bias=ln[2], ln=(ln[0], ln[1])) if ln is not None else {}
    dev = xq.device
    C = wo.shape[1] if wq is None else wq.shape[0]
    d = C // heads
    Lp = rup(L)
    if FUSED_ATTENTION and d % 8 == 0 and d <= 192:
        # Short key sequences with small heads (the 77-token context at d = 40) run the 32-key-step kernel, which takes V^T; every
        # other shape reads V in place ([B*L, C] slice of the projection output) as an MN-major tensor-core operand.
        in_place_v = not (L <= 128 and d <= 64)
        v = vt = None
        if kv is not None:
            q = gemm(xq, wq, **lnk)
            k, vt = kv
        elif wqkv is not None:
            qkv = gemm(xq, wqkv, **lnk)
            q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
        else:
            q = gemm(xq, wq)
            if wkv is not None and in_place_v:
                kvp = gemm(xkv, wkv)
                k, v = kvp[:, :C], kvp[:, C:]
            elif wkv is not None:
                k, vt = project_kv(xkv, B, L, wkv, heads)
            else:
                k, v = gemm(xkv, wk), gemm(xkv, wv)
        o = torch.empty((B * S, C), dtype=F16, device=dev)
        with torch.cuda.device(dev):
            if vt is None and in_place_v:
                call("coma_attention_fwd_nt_f16", q.data_ptr(), k.data_ptr(), v.data_ptr(), B, heads, S, L, d, q.stride(0), k.stride(0), v.stride(0),
                     float(d ** -0.5), o.data_ptr(), None, o.stride(0), _stream())
            else:
                if vt is None:
                    vt = torch.empty((B, heads, d, Lp), dtype=F16, device=dev)
                    call("coma_transpose_heads_f16", v.data_ptr(), B, L, heads, d, v.stride(0), vt.data_ptr(), Lp, _stream())
                call("coma_attention_fwd_f16", q.data_ptr(), k.data_ptr(), vt.data_ptr(), B, heads, S, L, d, q.stride(0), k.stride(0), Lp,
                     float(d ** -0.5), o.data_ptr(), o.stride(0), _stream())
        return gemm(o, wo, bo, residual, ln_out=ln_out)
    if wq is None:
        wq, wk, wv = wqkv[:C], wqkv[C:2 * C], wqkv[2 * C:]
    elif wk is None:
        wk, wv = wkv[:C], wkv[C:]
    q, k, v = gemm(xq, wq), gemm(xkv, wk), gemm(xkv, wv)
    scores = torch.empty((B, heads, S, Lp), dtype=F16, device=dev)
    gemm_batched(q, C, d, S * C, k, C, d, L * C, scores, Lp, S * Lp, heads * S * Lp, S, L, d, heads, B, alpha=d ** -0.5)
    vt = torch.empty((B, heads, d, Lp), dtype=F16, device=dev)
    with torch.cuda.device(dev):
        call("coma_softmax_rows_f16", scores.data_ptr(), B * heads * S, L, Lp, _stream())
        call("coma_transpose_heads_f16", v.data_ptr(), B, L, heads, d, v.stride(0), vt.data_ptr(), Lp, _stream())
    o = torch.empty((B * S, C), dtype=F16, device=dev)
    gemm_batched(scores, Lp, S * Lp, heads * S * Lp, vt, Lp, d * Lp, heads * d * Lp, o, C, d, S * C, S, d, Lp, heads, B)
    return gemm(o, wo, bo, residual)


# ------------------------------------------------------------------------------------------------ weight preparation
def prep_conv3x3(w, device):
    """[Cout, Cin, 3, 3] -> [Cout, rup(9*Cin)] f16 with K order (ky, kx, cin), zero padded."""
    cout, cin = w.shape[:2]
    k = w.permute(0, 2, 3, 1).reshape(cout, 9 * cin)
    out = torch.zeros((cout, rup(9 * cin)), dtype=F16, device=device)
    out[:, : 9 * cin] = k.to(device=device, dtype=F16)
    return out


def prep_linear(w, device):
    """[N, K] (or 1x1 conv [N, K, 1, 1]) -> [N, rup(K)] f16."""
    w = w.reshape(w.shape[0], -1)
    out = torch.zeros((w.shape[0], rup(w.shape[1])), dtype=F16, device=device)
    out[:, : w.shape[1]] = w.to(device=device, dtype=F16)
    return out


def prep_vec(v, device):
    return v.to(device=device, dtype=F32).contiguous()


class Graphed:
    """Captures `fn(*static_args)` into a CUDA graph (after an eager warm-up that also sizes the allocator pool and sets
    kernel attributes) and replays it: the ~650 launches of a UNet forward then cost one host call instead of being bound by
    per-launch host overhead. Inputs are updated IN PLACE in `static_args`; the output tensors are reused across replays."""

    def __init__(self, fn, *static_args):
        self.args = static_args
        fn(*static_args)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = fn(*static_args)

    def __call__(self):
        self.graph.replay()
        return self.out
