"""oracle/inpaint_loop_oracle.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

torch + cv2 restatement of the adaptive-mask denoising loop (utils/adaptive_mask_inpainting.py:908-1097, adapt_mask
:1123-1157, prepare_mask_and_masked_image :166-245, prepare_mask_latents :686-719) on top of oracle/sd_oracle.py.
cv2.dilate / np.logical_and here are the reference's own calls (true oracle for the mask logic); UNet / VAE / DDIM are
the UNPINNED restatements of sd_oracle. Random draws are injected by the caller so both sides consume identical noise."""
import cv2
import numpy as np
import torch
import torch.nn.functional as F

from . import sd_oracle as so


def adapt_mask_np(seg, default_u8, dilate_num, use_default_mask, human_detection_thres):
    """adapt_mask :1123-1141 for one image -> float32 mask in {0,1} [H,W]."""
    default = default_u8.astype(np.uint8) / 255                     # default_mask_image_np (:923)
    if use_default_mask or seg.sum() < 512 * 512 * human_detection_thres:
        mask = default
    else:
        mask = cv2.dilate(seg, np.ones((3, 3), dtype=np.uint8), iterations=dilate_num)
        mask = np.logical_and(mask, default)
    mask = torch.tensor(mask, dtype=torch.float32)
    mask[mask < 0.5] = 0                                             # :202-203
    mask[mask >= 0.5] = 1
    return mask.numpy()


def run_loop(unet_sd, vae_sd, ucfg, vcfg, image_u8, default_u8, ctx2, timesteps, ratio, guidance, strength, draws, segmenter,
             settings, human_detection_thres, enforce_full_mask_ratio, emulate_fp16=True, device="cpu"):
    """One work item (batch 1, like the reference). draws: list of [4,h,w] fp32 tensors consumed in the reference's RNG order.
    Returns per-step dicts (latents, x0, mask64) and the final image in [0,1]."""
    dev = torch.device(device)
    ac = so.ddim_alphas_cumprod().double()
    draws = [d.to(dev) for d in draws]
    img = (torch.from_numpy(image_u8.copy()).float() / 127.5 - 1.0).permute(2, 0, 1)[None].to(dev)
    H, W = image_u8.shape[:2]
    scal = vcfg["scaling_factor"]

    def enc_sample(x):
        m, lv = so.vae_encode_moments(vae_sd, x, vcfg, emulate_fp16)
        return (m + torch.exp(0.5 * lv) * draws.pop(0)[None]) * scal

    def mask_tensors(mask_np):
        mask = torch.from_numpy(mask_np)[None, None].to(dev)
        masked = img * (mask < 0.5)
        m64 = F.interpolate(mask, size=(H // 8, W // 8))
        return masked, m64

    mask_np = adapt_mask_np(np.zeros((H, W), np.uint8), default_u8, 0, True, human_detection_thres)
    masked, m64 = mask_tensors(mask_np)
    if strength == 1.0:
        latents = draws.pop(0)[None]
    else:
        il = enc_sample(img)
        noise = draws.pop(0)[None]
        a = float(ac[timesteps[0]])
        latents = a ** 0.5 * il + (1 - a) ** 0.5 * noise
    ml = enc_sample(masked.half().float() if emulate_fp16 else masked)
    out = []
    for i, t in enumerate(timesteps):
        x9 = torch.cat([latents, m64, ml], 1)
        x9 = torch.cat([x9, x9], 0)
        tt = torch.full((2,), float(t), device=dev)
        eps = so.unet_forward(unet_sd, x9, tt, ctx2, ucfg, emulate_fp16)
        e = eps[:1] + guidance * (eps[1:] - eps[:1])
        a_t = float(ac[t])
        a_p = float(ac[t - ratio] if t - ratio >= 0 else ac[0])
        x0 = (latents - (1 - a_t) ** 0.5 * e) / a_t ** 0.5
        latents = a_p ** 0.5 * x0 + (1 - a_p) ** 0.5 * e
        if enforce_full_mask_ratio > 0.0:
            use_default = t < 1000 * enforce_full_mask_ratio
        else:
            use_default = False
        if settings.provoke_scheduler(i):
            z = x0 / scal
            dec = so.vae_decode(vae_sd, z.half().float() if emulate_fp16 else z, vcfg, emulate_fp16)
            pred = ((dec / 2 + 0.5).clamp(0, 1)[0].permute(1, 2, 0).cpu().numpy() * 255).astype(np.uint8)
            seg = np.asarray(segmenter(pred)["mask"]).astype(np.uint8)
            mask_np = adapt_mask_np(seg, default_u8, settings.dilate_scheduler(i), use_default, human_detection_thres)
            masked, m64 = mask_tensors(mask_np)
            ml = enc_sample(masked.half().float() if emulate_fp16 else masked)
        out.append(dict(t=t, latents=latents.clone(), x0=x0.clone(), mask64=m64.clone(), mask=mask_np.copy()))
    z = latents / scal
    final = (so.vae_decode(vae_sd, z.half().float() if emulate_fp16 else z, vcfg, emulate_fp16) / 2 + 0.5).clamp(0, 1)
    return out, final


def time_reference_loop(device, steps=50, strength=0.98, guidance=11.0, seed=0, dtype=torch.float16, sdpa=False):
    """Times the REFERENCE-SHAPED loop (batch 1 per call, fp16 torch eager: cuDNN convolutions, unfused
    baddbmm+softmax+bmm attention as under torch 1.13, VAE decode on EVERY step (:1028), cv2.dilate on the host with the
    D2H/H2D round trips) with the full-size restated models and seeded random weights — the 'reference single-GPU PyTorch
    path' of BASELINE.md §4.3. Returns seconds per image."""
    import time
    from coma_b200.inpaint.pipeline import DDIMSchedule, default_adaptive_mask_settings
    from coma_b200.inpaint.segmenter import LuminanceSegmenter
    dev = torch.device(device)
    usd = {k: v.to(dev, dtype) for k, v in so.make_unet_state_dict(0).items()}
    vsd = {k: v.to(dev, dtype) for k, v in so.make_vae_state_dict(1).items()}
    rng = np.random.default_rng(0)
    image = rng.integers(0, 256, (512, 512, 3), dtype=np.uint8)
    default = np.zeros((512, 512), np.uint8)
    default[64:448, 128:384] = 255
    ts, ratio = DDIMSchedule().timesteps(steps, strength)
    settings = default_adaptive_mask_settings(steps)
    seg = LuminanceSegmenter(128)
    g = torch.Generator(device=dev).manual_seed(seed)
    ctx2 = (torch.randn((2, 77, 768), device=dev, generator=g) * 0.02).to(dtype)
    ac = so.ddim_alphas_cumprod().double()
    scal = so.VAE_CFG["scaling_factor"]

    def enc_sample(x):
        m, lv = so.vae_encode_moments(vsd, x, so.VAE_CFG)
        return (m + torch.exp(0.5 * lv) * torch.randn(m.shape, device=dev, generator=g, dtype=dtype)) * scal

    def one_image():
        img = (torch.from_numpy(image.copy()).float() / 127.5 - 1.0).permute(2, 0, 1)[None].to(dev, dtype)
        mask_np = adapt_mask_np(np.zeros((512, 512), np.uint8), default, 0, True, 0.015)
        mask = torch.from_numpy(mask_np)[None, None].to(dev, dtype)
        masked, m64 = img * (mask < 0.5), F.interpolate(mask, size=(64, 64))
        il = enc_sample(img)
        noise = torch.randn(il.shape, device=dev, generator=g, dtype=dtype)
        a = float(ac[ts[0]])
        latents = a ** 0.5 * il + (1 - a) ** 0.5 * noise
        ml = enc_sample(masked)
        for i, t in enumerate(ts):
            x9 = torch.cat([latents, m64, ml], 1)
            eps = so.unet_forward(usd, torch.cat([x9, x9], 0), torch.full((2,), float(t), device=dev), ctx2)
            e = eps[:1] + guidance * (eps[1:] - eps[:1])
            a_t, a_p = float(ac[t]), float(ac[t - ratio] if t - ratio >= 0 else ac[0])
            x0 = (latents - (1 - a_t) ** 0.5 * e) / a_t ** 0.5
            latents = a_p ** 0.5 * x0 + (1 - a_p) ** 0.5 * e
            dec = so.vae_decode(vsd, x0 / scal)                       # every step, like the reference
            pred = ((dec.float() / 2 + 0.5).clamp(0, 1)[0].permute(1, 2, 0).cpu().numpy() * 255).astype(np.uint8)
            if settings.provoke_scheduler(i):
                s = np.asarray(seg(pred)["mask"]).astype(np.uint8)
                mask_np = adapt_mask_np(s, default, settings.dilate_scheduler(i), False, 0.015)
                mask = torch.from_numpy(mask_np)[None, None].to(dev, dtype)
                masked, m64 = img * (mask < 0.5), F.interpolate(mask, size=(64, 64))
                ml = enc_sample(masked)
        out = so.vae_decode(vsd, latents / scal)
        return out

    so.USE_SDPA = bool(sdpa)   # True: F.scaled_dot_product_attention; False: torch-1.13-style unfused attention
    try:
        with torch.no_grad():
            one_image()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            one_image()
            torch.cuda.synchronize()
            return time.perf_counter() - t0
    finally:
        so.USE_SDPA = False
