"""AdaptiveMaskInpaintPipeline on B200 — drop-in mirror of the reference pipeline's call surface
(utils/adaptive_mask_inpainting.py:248, `__call__` :731-1109, `adapt_mask` :1123-1157, schedulers :1457-1485) with a batch
extension: B work items that share (render, default mask, prompt) and differ by seed run together, each with its own
adaptive mask and masked-image latents (UNet batch 2B).

Differences from the reference, all output-preserving:
  * x0 is decoded only on provoke steps (the reference decodes every step, :1028, but only consumes the image there);
  * mask logic (area test, dilation, AND, binarise, masked image, /8 mask) is one GPU call instead of cv2 on the host;
  * CFG + DDIM run as one fused fp32 kernel; latents are kept in fp32 between steps.
"""
import os
from dataclasses import dataclass, field

import numpy as np
import torch

from . import nn
from .nn import Act, F16, F32
from .._lib import call, _stream


class MaskDilateScheduler:
    """utils/adaptive_mask_inpainting.py:1457-1465."""

    def __init__(self, max_dilate_num=15, num_inference_steps=50, schedule=None):
        self.max_dilate_num = max_dilate_num
        self.schedule = [num_inference_steps - i for i in range(num_inference_steps)] if schedule is None else schedule
        assert len(self.schedule) == num_inference_steps

    def __call__(self, i):
        return min(self.max_dilate_num, self.schedule[i])


class ProvokeScheduler:
    """utils/adaptive_mask_inpainting.py:1468-1485 (1-indexed unless is_zero_indexing)."""

    def __init__(self, num_inference_steps=50, schedule=None, is_zero_indexing=False):
        schedule = [] if schedule is None else schedule
        if len(schedule) > 0:
            assert max(schedule) <= (num_inference_steps - 1 if is_zero_indexing else num_inference_steps)
        self.is_zero_indexing, self.schedule = is_zero_indexing, schedule

    def __call__(self, i):
        return (i if self.is_zero_indexing else i + 1) in self.schedule


@dataclass
class AdaptiveMaskSettings:
    dilate_scheduler: MaskDilateScheduler
    provoke_scheduler: ProvokeScheduler
    dilate_kernel: np.ndarray = field(default_factory=lambda: np.ones((3, 3), dtype=np.uint8))


def default_adaptive_mask_settings(ddim_steps=50):
    """The settings src/generation/inpaint.py:112-132 registers for the PointRend segmenter (type "p")."""
    n = int(ddim_steps * 0.1)
    sched = [20] * n + [10] * n + [5] * n + [4] * n + [3] * n + [2] * n + [1] * n + [0] * (ddim_steps - 7 * n)
    return AdaptiveMaskSettings(MaskDilateScheduler(20, ddim_steps, sched),
                                ProvokeScheduler(ddim_steps, list(range(2, 11, 2)) + list(range(12, 41, 2)) + [45], False))


class DDIMSchedule:
    """DDIMScheduler as built at src/generation/inpaint.py:54-60 (scaled_linear betas 0.00085->0.012, 1000 train steps,
    clip_sample False, set_alpha_to_one False, steps_offset forced to 1 by the pipeline's __init__ :295-307)."""

    def __init__(self, num_train=1000, beta_start=0.00085, beta_end=0.012):
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0).double().numpy()
        self.num_train = num_train

    def timesteps(self, num_inference_steps, strength=1.0):
        ratio = self.num_train // num_inference_steps
        ts = [i * ratio + 1 for i in range(num_inference_steps)][::-1]
        init = min(int(num_inference_steps * strength), num_inference_steps)   # get_timesteps :722-729
        return ts[max(num_inference_steps - init, 0):], ratio

    def alphas(self, t, ratio):
        prev = t - ratio
        return float(self.alphas_cumprod[t]), float(self.alphas_cumprod[prev] if prev >= 0 else self.alphas_cumprod[0])


@dataclass
class PipelineOutput:
    images: list
    nsfw_content_detected: object = None
    masks: object = None       # final adaptive masks, u8 [B,H,W] (cuda)
    trace: object = None       # optional per-step tensors (tests)


class AdaptiveMaskInpaintPipeline:
    vae_scale_factor = 8

    def __init__(self, unet, vae, text_encoder=None, scheduler=None, use_cuda_graphs=True):
        self.unet, self.vae, self.text_encoder = unet, vae, text_encoder
        self.use_cuda_graphs = use_cuda_graphs
        self._graphs = {}
        self.scheduler = scheduler or DDIMSchedule()
        self.dev = unet.dev
        self.adaptive_mask_model = None
        self.adaptive_mask_settings = None

    def register_adaptive_mask_settings(self, adaptive_mask_settings):
        self.adaptive_mask_settings = adaptive_mask_settings

    def register_adaptive_mask_model(self, adaptive_mask_model):
        self.adaptive_mask_model = adaptive_mask_model

    # ------------------------------------------------------------------------------------------------ helpers
    def _embeds(self, prompt, negative_prompt, prompt_embeds, negative_prompt_embeds, B, L):
        if prompt_embeds is None:
            if self.text_encoder is None:
                raise ValueError("no text encoder is registered (CLIP weights are not shipped): pass `prompt_embeds` / "
                                 "`negative_prompt_embeds` [77, cross_dim]")
            prompt_embeds = self.text_encoder(prompt)
            negative_prompt_embeds = self.text_encoder(negative_prompt or "")
        elif negative_prompt_embeds is None:
            if self.text_encoder is None:
                raise ValueError("`prompt_embeds` was given without `negative_prompt_embeds`: classifier-free guidance needs both "
                                 "(or a registered text encoder to embed `negative_prompt`)")
            negative_prompt_embeds = self.text_encoder(negative_prompt or "")
        pe = torch.as_tensor(prompt_embeds, device=self.dev).to(F16).reshape(1, L, -1)
        ne = torch.as_tensor(negative_prompt_embeds, device=self.dev).to(F16).reshape(1, L, -1)
        return torch.cat([ne.expand(B, -1, -1), pe.expand(B, -1, -1)], 0).reshape(2 * B * L, -1).contiguous()  # :536-552 order

    def _randn(self, shape, generators):
        """One draw per work item from its own generator, in the reference's NCHW fp16 shape (diffusers randn_tensor)."""
        outs = [torch.randn((1,) + shape, generator=g, device=self.dev, dtype=F16) for g in generators]
        return torch.cat(outs, 0).float().permute(0, 2, 3, 1).reshape(-1, shape[0]).contiguous()     # -> [B*h*w, C] fp32

    def _graphed(self, key, fn, *static_args):
        """Run fn(*static_args) through a cached CUDA graph (captured on first use for this key)."""
        if not self.use_cuda_graphs:
            return fn(*static_args)
        if key not in self._graphs:
            self._graphs[key] = nn.Graphed(fn, *static_args)
        return self._graphs[key]()

    def _static(self, key, make):
        buf = self._graphs.get(("buf",) + key)
        if buf is None:
            buf = self._graphs[("buf",) + key] = make()
        return buf

    def _encode_moments(self, img_act: Act):
        key = ("enc", img_act.B, img_act.H, img_act.W)
        static = self._static(key, lambda: nn.new_act(img_act.B, img_act.H, img_act.W, 3, self.dev))
        static.t.copy_(img_act.t)
        return self._graphed(key, self.vae.encode_moments, static)

    def _decode(self, z: Act):
        key = ("dec", z.B, z.H, z.W)
        static = self._static(key, lambda: nn.new_act(z.B, z.H, z.W, z.C, self.dev))
        static.t.copy_(z.t)
        return self._graphed(key, lambda a: self.vae.decode(a, out_dtype=F32), static)

    def _encode_sample(self, img_act: Act, generators):
        mean, logvar = self._encode_moments(img_act)
        h, w = img_act.H // 8, img_act.W // 8
        noise = self._randn((mean.shape[1], h, w), generators)
        out = torch.empty_like(mean)
        with torch.cuda.device(self.dev):
            call("coma_sample_latents_f32", mean.data_ptr(), logvar.data_ptr(), noise.data_ptr(), mean.numel(),
                 float(self.vae.cfg["scaling_factor"]), out.data_ptr(), _stream())
        return out

    def _mask_update(self, seg_u8, default_u8, image_f32, B, H, W, dilate, area_thres, force_default):
        dev = self.dev
        scratch = torch.empty(2 * B * H * W, dtype=torch.uint8, device=dev)
        mask = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
        masked = nn.new_act(B, H, W, 3, dev)
        small = torch.empty((B * (H // 8) * (W // 8),), dtype=F32, device=dev)
        used = torch.empty(B, dtype=torch.int32, device=dev)
        area = torch.empty(B, dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            call("coma_adaptive_mask_u8", seg_u8.data_ptr(), default_u8.data_ptr(), B, H, W, int(dilate), float(area_thres),
                 int(force_default), image_f32.data_ptr(), scratch.data_ptr(), mask.data_ptr(), masked.t.data_ptr(), masked.ld,
                 small.data_ptr(), used.data_ptr(), area.data_ptr(), _stream())
        return mask, masked, small, used

    def decode_to_uint8(self, latents, B, h, w):
        """decode_to_npuint8_image (:1111-1115) for the whole batch -> uint8 cuda tensor [B,8h,8w,3]."""
        z = nn.new_act(B, h, w, latents.shape[1], self.dev)
        z.t.copy_(latents / self.vae.cfg["scaling_factor"])
        img = self._decode(z)
        out = torch.empty((B, img.H, img.W, 3), dtype=torch.uint8, device=self.dev)
        with torch.cuda.device(self.dev):
            call("coma_image_to_u8", img.t.data_ptr(), img.M, img.ld, out.data_ptr(), _stream())
        return out

    # ------------------------------------------------------------------------------------------------ the loop
    @torch.no_grad()
    def __call__(self, prompt=None, image=None, default_mask_image=None, negative_prompt=None, height=None, width=None,
                 strength=1.0, num_inference_steps=50, guidance_scale=7.5, use_adaptive_mask=True, generator=None,
                 enforce_full_mask_ratio=0.5, human_detection_thres=0.008, visualization_save_dir=None, prompt_embeds=None,
                 negative_prompt_embeds=None, batch_size=1, output_type="pil", return_trace=False, **_ignored):
        if image is None:
            raise ValueError("`image` input cannot be undefined.")
        if default_mask_image is None:
            raise ValueError("`mask_image` input cannot be undefined.")
        if strength < 0 or strength > 1:
            raise ValueError(f"The value of strength should in [0.0, 1.0] but is {strength}")
        if self.unet.cfg["in_channels"] != 9:
            raise NotImplementedError  # :998
        if use_adaptive_mask and (self.adaptive_mask_model is None or self.adaptive_mask_settings is None):
            raise ValueError("register_adaptive_mask_model / register_adaptive_mask_settings must be called first")
        dev, B = self.dev, int(batch_size)
        img_np = np.asarray(image.convert("RGB") if hasattr(image, "convert") else image)
        H, W = img_np.shape[:2]
        if H % 8 or W % 8:
            raise ValueError(f"`height` and `width` have to be divisible by 8 but are {H} and {W}.")
        h, w = H // 8, W // 8
        if hasattr(default_mask_image, "convert"):
            default_mask_image = default_mask_image.convert("L")
            if default_mask_image.size != (W, H):   # the reference resizes the PIL mask with the image (prepare_mask_and_masked_image)
                default_mask_image = default_mask_image.resize((W, H))
        default_np = np.asarray(default_mask_image).astype(np.uint8)
        if default_np.shape != (H, W):
            raise ValueError(f"`default_mask_image` must be a single-channel [H, W] = [{H}, {W}] mask matching `image`, got {default_np.shape}")
        gens = generator if isinstance(generator, (list, tuple)) else [generator] * B
        if len(gens) != B:
            raise ValueError(f"You have passed a list of generators of length {len(gens)}, but requested an effective batch size of {B}.")
        L = 77
        ctx = self._embeds(prompt, negative_prompt, prompt_embeds, negative_prompt_embeds, B, L)
        do_cfg = guidance_scale > 1.0
        assert do_cfg, "the B200 loop is built for classifier-free guidance (guidance_scale > 1), as every reference config uses"

        timesteps, ratio = self.scheduler.timesteps(num_inference_steps, strength)
        if len(timesteps) < 1:
            raise ValueError(f"After adjusting the num_inference_steps by strength parameter: {strength}, the number of pipeline steps is {len(timesteps)} which is < 1 and not appropriate for this pipeline.")

        # image / default mask on the device; initial mask = binarised default (:922 prepare_mask_and_masked_image)
        image_f32 = (torch.from_numpy(img_np.copy()).to(dev).float() / 127.5 - 1.0).unsqueeze(0).expand(B, -1, -1, -1).contiguous()
        default_u8 = torch.from_numpy(default_np.copy()).to(dev)
        zero_seg = torch.zeros((B, H, W), dtype=torch.uint8, device=dev)
        area_thres = 512 * 512 * human_detection_thres                      # hard-coded 512^2, :1130
        mask_u8, masked_img, mask64, _ = self._mask_update(zero_seg, default_u8, image_f32, B, H, W, 0, area_thres, True)

        # latents (:931 prepare_latents): RNG order = encode(image).sample -> noise -> encode(masked).sample -> one per adapt
        img_act = nn.new_act(B, H, W, 3, dev)
        img_act.t.copy_(image_f32.reshape(-1, 3))
        is_strength_max = strength == 1.0
        image_latents = None if is_strength_max else self._encode_sample(img_act, gens)
        noise = self._randn((4, h, w), gens)
        if is_strength_max:
            latents = noise.clone()                                            # init_noise_sigma = 1 for DDIM
        else:
            a = float(self.scheduler.alphas_cumprod[timesteps[0]])
            latents = (a ** 0.5) * image_latents + ((1 - a) ** 0.5) * noise     # scheduler.add_noise at the first timestep
        masked_latents = self._encode_sample(masked_img, gens)                  # :953 prepare_mask_latents

        rows = B * h * w
        ukey = ("unet", B, h, w, L)
        x_in = self._static(ukey + ("x",), lambda: nn.new_act(2 * B, h, w, 9, dev))
        tt = self._static(ukey + ("t",), lambda: torch.zeros((2 * B,), dtype=F32, device=dev))
        ctx_static = self._static(ukey + ("ctx",), lambda: torch.empty_like(ctx))
        ctx_static.copy_(ctx)
        # cross-attention keys / values depend on the prompt only: projected once per call, reused by all UNet evaluations
        ctx_kv = self._graphed(ukey + ("kv",), lambda c: self.unet.context_kv(c, L, 2 * B), ctx_static)
        trace = [] if return_trace else None
        settings = self.adaptive_mask_settings
        for i, t in enumerate(timesteps):
            with torch.cuda.device(dev):
                call("coma_assemble_unet_input_f16", latents.data_ptr(), mask64.data_ptr(), masked_latents.data_ptr(), rows,
                     x_in.t.data_ptr(), x_in.ld, _stream())
            tt.fill_(float(t))
            eps = self._graphed(ukey, lambda a, b, c: self.unet.forward(a, b, c, L, ctx_kv=ctx_kv), x_in, tt, ctx_static)  # fp32 [2*rows, 4 (ld 8)]
            a_t, a_prev = self.scheduler.alphas(t, ratio)
            new_latents, x0 = torch.empty_like(latents), torch.empty_like(latents)
            with torch.cuda.device(dev):
                call("coma_cfg_ddim_step_f32", eps.data_ptr(), rows, eps.stride(0), 4, float(guidance_scale), latents.data_ptr(),
                     a_t, a_prev, new_latents.data_ptr(), x0.data_ptr(), _stream())
            latents = new_latents
            if use_adaptive_mask:
                if enforce_full_mask_ratio > 0.0:
                    use_default = t < self.scheduler.num_train * enforce_full_mask_ratio
                elif enforce_full_mask_ratio == 0.0:
                    use_default = False
                else:
                    raise NotImplementedError
                if settings.provoke_scheduler(i):
                    pred = self.decode_to_uint8(x0, B, h, w)                      # [B,H,W,3] u8 on the device
                    seg = self._segment(pred)
                    mask_u8, masked_img, mask64, _ = self._mask_update(seg, default_u8, image_f32, B, H, W,
                                                                       settings.dilate_scheduler(i), area_thres, use_default)
                    masked_latents = self._encode_sample(masked_img, gens)
                    if visualization_save_dir is not None and getattr(self.adaptive_mask_model, "use_visualizer", False):
                        self._dump_step(visualization_save_dir, i, mask_u8, pred)
            if trace is not None:
                trace.append(dict(t=t, latents=latents.clone(), x0=x0.clone(), mask64=mask64.clone()))

        final = self.decode_to_uint8_final(latents, B, h, w, output_type)
        return PipelineOutput(images=final, masks=mask_u8, trace=trace)

    def _dump_step(self, save_dir, i, mask_u8, pred_u8):
        """Per-provoke-step dump of the adapted mask and the decoded x0 (utils/adaptive_mask_inpainting.py:1051-1060: `masks/<i:05>.png`
        with the reference's grey ramp clip(0.6 + (1 - mask), 0, 1), `images/<i:05>.png`); batch element b > 0 goes to `<save_dir>/<b>/`.
        A debugging aid: it synchronises and copies to the host, exactly like the reference's visualiser."""
        from PIL import Image
        masks, imgs = mask_u8.cpu().numpy(), pred_u8.cpu().numpy()
        for b in range(masks.shape[0]):
            root = save_dir if b == 0 else os.path.join(save_dir, str(b))
            os.makedirs(os.path.join(root, "masks"), exist_ok=True)
            os.makedirs(os.path.join(root, "images"), exist_ok=True)
            m = (masks[b] > 0).astype(np.float32)
            Image.fromarray((np.clip(0.6 + (1.0 - m), 0.0, 1.0) * 255).astype(np.uint8)).convert("L").save(os.path.join(root, "masks", f"{i:05}.png"))
            Image.fromarray(imgs[b]).save(os.path.join(root, "images", f"{i:05}.png"))

    def _segment(self, pred_u8):
        m = self.adaptive_mask_model
        if getattr(m, "accepts_cuda", False):
            return m(pred_u8)["mask"].to(torch.uint8).contiguous()
        host = pred_u8.cpu().numpy()                                              # the reference's D2H round trip (:1114)
        masks = np.stack([np.asarray(m(host[b])["mask"]).astype(np.uint8) for b in range(host.shape[0])])
        return torch.from_numpy(masks).to(self.dev)

    def decode_to_uint8_final(self, latents, B, h, w, output_type):
        """:1086-1097: final decode + VaeImageProcessor.postprocess (denormalise, clamp, round to uint8 like PIL does)."""
        z = nn.new_act(B, h, w, latents.shape[1], self.dev)
        z.t.copy_(latents / self.vae.cfg["scaling_factor"])
        img = self._decode(z)
        x = (img.t[:, :3].reshape(B, img.H, img.W, 3) * 0.5 + 0.5).clamp(0, 1)
        if output_type == "pt":
            return x
        arr = (x * 255).round().to(torch.uint8).cpu().numpy()                     # numpy_to_pil: (images*255).round()
        if output_type == "np":
            return arr
        from PIL import Image
        return [Image.fromarray(a) for a in arr]
