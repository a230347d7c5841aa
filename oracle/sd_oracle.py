"""oracle/sd_oracle.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Plain-torch fp32 restatement of the models the adaptive-mask inpainting loop calls
(utils/adaptive_mask_inpainting.py:1001-1007 `self.unet`, :680 `self.vae.encode`, :1086/:1112 `self.vae.decode`,
:1015 `self.scheduler.step`). Their arithmetic lives in diffusers==0.20.2 (INSTALL.md:31), which is NOT vendored in the
reference and not installed here, and no weights are available: this file restates the published architecture of
`UNet2DConditionModel` (SD-1.5 inpainting config: in 9 / out 4, blocks (320,640,1280,1280), 2 layers per block, 8 heads,
cross-attention dim 768, GroupNorm(32)), `AutoencoderKL` (blocks (128,256,512,512), latent 4, scaling 0.18215) and
`DDIMScheduler.step` (epsilon prediction, eta = 0) with diffusers' state-dict key names, so real checkpoints can be
dropped in later.  **Parity status: UNPINNED** — there is no diffusers install, golden vector or reference test to check
this restatement against; only `cv2.dilate` (mask logic) is a true oracle on this path.

`emulate_fp16=True` rounds activations to fp16 at the points where the B200 path stores fp16 (layer outputs), so that
block-level comparisons isolate accumulation-order differences from storage-precision differences.
"""
import math

import torch
import torch.nn.functional as F

# configurations + seeded random state dicts: shared with the product-side benchmark (a data generator, not an algorithm)
from coma_b200.inpaint.synthetic import (UNET_CFG, VAE_CFG, make_unet_state_dict, make_vae_state_dict, tiny_unet_cfg,  # noqa: E402,F401
                                          tiny_vae_cfg)

USE_SDPA = False   # set by inpaint_loop_oracle.time_reference_loop(sdpa=True)


def round_weights_fp16(sd):
    """The models run with fp16 weights (src/generation/inpaint.py:64 torch_dtype=float16): both sides of every parity
    test use these fp16-rounded values (norm affine / biases stay fp32 in the B200 path and are rounded here too)."""
    return {k: v.half().float() for k, v in sd.items()}


# ------------------------------------------------------------------------------------------------ forward passes
class _R:
    """Optional fp16 rounding of stored activations."""

    def __init__(self, on):
        self.on = on

    def __call__(self, x):
        return x.half().float() if self.on else x


def _gn(x, sd, name, groups, eps):
    return F.group_norm(x, groups, sd[name + ".weight"], sd[name + ".bias"], eps)


def _conv(x, sd, name, stride=1, padding=1):
    w = sd[name + ".weight"]
    return F.conv2d(x.to(w.dtype), w, sd[name + ".bias"], stride=stride, padding=padding)


def _lin(x, sd, name):
    w = sd[name + ".weight"]
    return F.linear(x.to(w.dtype), w, sd.get(name + ".bias"))


def resnet_block(x, temb, sd, name, groups, eps, r):
    """diffusers ResnetBlock2D (time_embedding_norm='default', output_scale_factor 1). The B200 path applies GN+SiLU while
    building the conv operand in fp16, hence the rounding of the conv inputs."""
    h = r(F.silu(_gn(x, sd, name + ".norm1", groups, eps)))
    h = _conv(h, sd, name + ".conv1")
    if temb is not None:
        h = h + _lin(r(F.silu(temb)), sd, name + ".time_emb_proj")[:, :, None, None]
    h = r(h)
    h = r(F.silu(_gn(h, sd, name + ".norm2", groups, eps)))
    h = _conv(h, sd, name + ".conv2")
    sc = _conv(x, sd, name + ".conv_shortcut", padding=0) if name + ".conv_shortcut.weight" in sd else x
    if name + ".conv_shortcut.weight" in sd:
        sc = r(sc)
    return r(sc + h)


def attention(xq, xkv, sd, name, heads, r, bias_out=True):
    B, S, C = xq.shape
    q, k, v = r(_lin(xq, sd, name + ".to_q")), r(_lin(xkv, sd, name + ".to_k")), r(_lin(xkv, sd, name + ".to_v"))
    d = C // heads
    sp = lambda t: t.view(B, -1, heads, d).permute(0, 2, 1, 3)
    if USE_SDPA and not r.on:   # timing baseline only (bench.py hoi reference arm, "sdpa" variant)
        o = F.scaled_dot_product_attention(sp(q), sp(k), sp(v)).permute(0, 2, 1, 3).reshape(B, S, C)
        return _lin(o, sd, name + ".to_out.0")
    s = r(torch.matmul(sp(q), sp(k).transpose(-1, -2)) * d ** -0.5)
    p = r(torch.softmax(s, dim=-1))
    o = r(torch.matmul(p, sp(v)).permute(0, 2, 1, 3).reshape(B, S, C))
    return _lin(o, sd, name + ".to_out.0")


def transformer_2d(x, ctx, sd, name, heads, groups, r):
    """diffusers Transformer2DModel (use_linear_projection=False) + BasicTransformerBlock (GEGLU feed-forward)."""
    B, C, H, W = x.shape
    res = x
    h = r(_gn(x, sd, name + ".norm", groups, 1e-6))
    h = r(_conv(h, sd, name + ".proj_in", padding=0))
    h = h.permute(0, 2, 3, 1).reshape(B, H * W, C)
    t = name + ".transformer_blocks.0"
    ln = lambda z, n: r(F.layer_norm(z, (C,), sd[n + ".weight"], sd[n + ".bias"], 1e-5))
    n1 = ln(h, t + ".norm1")
    h = r(attention(n1, n1, sd, t + ".attn1", heads, r) + h)
    h = r(attention(ln(h, t + ".norm2"), ctx, sd, t + ".attn2", heads, r) + h)
    ff = r(_lin(ln(h, t + ".norm3"), sd, t + ".ff.net.0.proj"))
    a, g = ff.chunk(2, dim=-1)
    ff = r(a * F.gelu(g))
    h = r(_lin(ff, sd, t + ".ff.net.2") + h)
    h = h.reshape(B, H, W, C).permute(0, 3, 1, 2)
    return r(_conv(h, sd, name + ".proj_out", padding=0) + res)


def timestep_embedding(t, dim):
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    a = t.float()[:, None] * freqs[None]
    return torch.cat([torch.cos(a), torch.sin(a)], dim=-1)  # flip_sin_to_cos=True


def unet_forward(sd, x, t, ctx, cfg=UNET_CFG, emulate_fp16=False, return_taps=False):
    """x [B,9,h,w], t [B] (float timesteps), ctx [B,77,cross_dim] -> eps [B,4,h,w] (fp32)."""
    r = _R(emulate_fp16)
    ch, L, heads, G = cfg["block_out_channels"], cfg["layers_per_block"], cfg["heads"], cfg["groups"]
    taps = {}
    x, ctx = r(x), r(ctx)
    temb = r(timestep_embedding(t, ch[0]))
    temb = r(F.silu(_lin(temb, sd, "time_embedding.linear_1")))
    temb = r(_lin(temb, sd, "time_embedding.linear_2"))
    h = r(_conv(x, sd, "conv_in"))
    skips = [h]
    for i in range(len(ch)):
        for j in range(L):
            h = resnet_block(h, temb, sd, f"down_blocks.{i}.resnets.{j}", G, 1e-5, r)
            if cfg["attn_levels"][i]:
                h = transformer_2d(h, ctx, sd, f"down_blocks.{i}.attentions.{j}", heads, G, r)
            skips.append(h)
        if i < len(ch) - 1:
            h = r(_conv(h, sd, f"down_blocks.{i}.downsamplers.0.conv", stride=2))
            skips.append(h)
    taps["down"] = h
    h = resnet_block(h, temb, sd, "mid_block.resnets.0", G, 1e-5, r)
    h = transformer_2d(h, ctx, sd, "mid_block.attentions.0", heads, G, r)
    h = resnet_block(h, temb, sd, "mid_block.resnets.1", G, 1e-5, r)
    taps["mid"] = h
    ral = list(reversed(cfg["attn_levels"]))
    for i in range(len(ch)):
        for j in range(L + 1):
            h = torch.cat([h, skips.pop()], dim=1)
            h = resnet_block(h, temb, sd, f"up_blocks.{i}.resnets.{j}", G, 1e-5, r)
            if ral[i]:
                h = transformer_2d(h, ctx, sd, f"up_blocks.{i}.attentions.{j}", heads, G, r)
        if i < len(ch) - 1:
            h = r(_conv(F.interpolate(h, scale_factor=2.0, mode="nearest"), sd, f"up_blocks.{i}.upsamplers.0.conv"))
    taps["up"] = h
    h = r(F.silu(_gn(h, sd, "conv_norm_out", G, 1e-5)))
    out = _conv(h, sd, "conv_out")
    return (out, taps) if return_taps else out


def _vae_attn(x, sd, name, G, r):
    B, C, H, W = x.shape
    h = r(_gn(x, sd, name + ".group_norm", G, 1e-6)).permute(0, 2, 3, 1).reshape(B, H * W, C)
    o = attention(h, h, sd, name, 1, r)
    return r(o.reshape(B, H, W, C).permute(0, 3, 1, 2) + x)


def vae_decode(sd, z, cfg=VAE_CFG, emulate_fp16=False):
    """z [B,4,h,w] (already divided by scaling_factor) -> image [B,3,8h,8w] in [-1,1] (AutoencoderKL.decode)."""
    r = _R(emulate_fp16)
    ch, L, G = cfg["block_out_channels"], cfg["layers_per_block"], cfg["groups"]
    h = r(_conv(r(z), sd, "post_quant_conv", padding=0))
    h = r(_conv(h, sd, "decoder.conv_in"))
    h = resnet_block(h, None, sd, "decoder.mid_block.resnets.0", G, 1e-6, r)
    h = _vae_attn(h, sd, "decoder.mid_block.attentions.0", G, r)
    h = resnet_block(h, None, sd, "decoder.mid_block.resnets.1", G, 1e-6, r)
    for i in range(len(ch)):
        for j in range(L + 1):
            h = resnet_block(h, None, sd, f"decoder.up_blocks.{i}.resnets.{j}", G, 1e-6, r)
        if i < len(ch) - 1:
            h = r(_conv(F.interpolate(h, scale_factor=2.0, mode="nearest"), sd, f"decoder.up_blocks.{i}.upsamplers.0.conv"))
    h = r(F.silu(_gn(h, sd, "decoder.conv_norm_out", G, 1e-6)))
    return _conv(h, sd, "decoder.conv_out")


def vae_encode_moments(sd, img, cfg=VAE_CFG, emulate_fp16=False):
    """img [B,3,H,W] in [-1,1] -> (mean, logvar) [B,4,H/8,W/8] (AutoencoderKL.encode -> DiagonalGaussianDistribution)."""
    r = _R(emulate_fp16)
    ch, L, G = cfg["block_out_channels"], cfg["layers_per_block"], cfg["groups"]
    h = r(_conv(r(img), sd, "encoder.conv_in"))
    for i in range(len(ch)):
        for j in range(L):
            h = resnet_block(h, None, sd, f"encoder.down_blocks.{i}.resnets.{j}", G, 1e-6, r)
        if i < len(ch) - 1:
            h = r(_conv(F.pad(h, (0, 1, 0, 1)), sd, f"encoder.down_blocks.{i}.downsamplers.0.conv", stride=2, padding=0))
    h = resnet_block(h, None, sd, "encoder.mid_block.resnets.0", G, 1e-6, r)
    h = _vae_attn(h, sd, "encoder.mid_block.attentions.0", G, r)
    h = resnet_block(h, None, sd, "encoder.mid_block.resnets.1", G, 1e-6, r)
    h = r(F.silu(_gn(h, sd, "encoder.conv_norm_out", G, 1e-6)))
    h = r(_conv(h, sd, "encoder.conv_out"))
    m = _conv(h, sd, "quant_conv", padding=0)
    mean, logvar = m.chunk(2, dim=1)
    return mean, torch.clamp(logvar, -30.0, 20.0)


# ------------------------------------------------------------------------------------------------ scheduler
def ddim_alphas_cumprod(num_train=1000, beta_start=0.00085, beta_end=0.012):
    """scaled_linear betas (src/generation/inpaint.py:54-60)."""
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


def ddim_timesteps(num_inference_steps=50, num_train=1000, steps_offset=1):
    """'leading' spacing + offset 1: 981, 961, ..., 1."""
    ratio = num_train // num_inference_steps
    return [int(i * ratio) + steps_offset for i in range(num_inference_steps)][::-1]


def ddim_step(eps, t, x, alphas_cumprod, num_inference_steps=50, num_train=1000):
    """DDIMScheduler.step (eta=0, epsilon prediction, clip_sample=False, set_alpha_to_one=False) -> (prev, x0)."""
    prev_t = t - num_train // num_inference_steps
    a_t = alphas_cumprod[t]
    a_prev = alphas_cumprod[prev_t] if prev_t >= 0 else alphas_cumprod[0]
    x0 = (x - (1 - a_t) ** 0.5 * eps) / a_t ** 0.5
    return a_prev ** 0.5 * x0 + (1 - a_prev) ** 0.5 * eps, x0
