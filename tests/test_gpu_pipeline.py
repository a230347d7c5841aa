"""HP-A loop parity on the GPU: mask logic (bit-exact vs cv2), CFG+DDIM step, and the whole adaptive-mask loop with toy-width
models against the torch+cv2 restatement (oracle/inpaint_loop_oracle.py)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


@pytest.mark.parametrize("k,force,thres", [(0, False, 0.001), (1, False, 0.001), (5, False, 0.001), (20, False, 0.001),
                                           (3, True, 0.001), (3, False, 0.9)])
def test_adaptive_mask_bit_exact_vs_cv2(dev, k, force, thres):
    """cv2.dilate(iterations=k) + logical_and + binarise + masked image + nearest /8 (utils/adaptive_mask_inpainting.py:1123-1141)."""
    from coma_b200.inpaint import nn
    from coma_b200._lib import call, _stream
    from oracle.inpaint_loop_oracle import adapt_mask_np
    rng = np.random.default_rng(k)
    B, H, W = 3, 512, 512
    seg = (rng.random((B, H, W)) < 0.002).astype(np.uint8)
    seg[1, 100:180, 200:260] = 1
    seg[2] = 0                                              # empty -> falls back to the default mask
    seg[0, 0, 0] = seg[0, H - 1, W - 1] = 1                  # borders
    default = np.zeros((H, W), np.uint8)
    default[64:448, 128:384] = 255
    default[10:20, 10:20] = 100                              # gray: truthy for logical_and, 0 after binarisation of the default
    image = rng.uniform(-1, 1, (B, H, W, 3)).astype(np.float32)
    t = lambda a: torch.from_numpy(a).to(dev)
    scratch = torch.empty(2 * B * H * W, dtype=torch.uint8, device=dev)
    mask = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
    masked = nn.new_act(B, H, W, 3, dev)
    small = torch.empty((B, H // 8, W // 8), dtype=torch.float32, device=dev)
    used = torch.empty(B, dtype=torch.int32, device=dev)
    area = torch.empty(B, dtype=torch.int64, device=dev)
    seg_d, default_d, image_d = t(seg), t(default), t(image)      # keep the device copies alive across the call
    call("coma_adaptive_mask_u8", seg_d.data_ptr(), default_d.data_ptr(), B, H, W, k, 512 * 512 * thres, int(force), image_d.data_ptr(),
         scratch.data_ptr(), mask.data_ptr(), masked.t.data_ptr(), masked.ld, small.data_ptr(), used.data_ptr(), area.data_ptr(), _stream())
    for b in range(B):
        ref = adapt_mask_np(seg[b], default, k, force, thres)
        np.testing.assert_array_equal(mask[b].cpu().numpy(), ref.astype(np.uint8))
        np.testing.assert_array_equal(small[b].cpu().numpy(), ref[::8, ::8])
        ref_masked = (image[b] * (ref[..., None] < 0.5)).astype(np.float16)
        np.testing.assert_array_equal(masked.t.reshape(B, H, W, 3)[b].cpu().numpy(), ref_masked)
        assert bool(used[b].item()) == (force or seg[b].sum() < 512 * 512 * thres)


def test_cfg_ddim_step(dev):
    from coma_b200._lib import call, _stream
    from oracle import sd_oracle as so
    g = torch.Generator(device=dev).manual_seed(0)
    rows = 2 * 64 * 64
    eps = torch.randn((2 * rows, 8), device=dev, generator=g)
    x = torch.randn((rows, 4), device=dev, generator=g)
    ac = so.ddim_alphas_cumprod().double()
    for t in (961, 21, 1):
        xp, x0 = torch.empty_like(x), torch.empty_like(x)
        a_prev = ac[t - 20] if t - 20 >= 0 else ac[0]
        call("coma_cfg_ddim_step_f32", eps.data_ptr(), rows, 8, 4, 11.0, x.data_ptr(), float(ac[t]), float(a_prev), xp.data_ptr(), x0.data_ptr(), _stream())
        e = eps[:rows, :4] + 11.0 * (eps[rows:, :4] - eps[:rows, :4])
        rp, r0 = so.ddim_step(e.double(), t, x.double(), ac.to(dev))
        torch.testing.assert_close(xp.double(), rp, rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(x0.double(), r0, rtol=1e-5, atol=1e-5)


def test_schedules_match_reference_constants():
    from coma_b200.inpaint.pipeline import DDIMSchedule, default_adaptive_mask_settings
    s = DDIMSchedule()
    ts, ratio = s.timesteps(50, 0.98)
    assert ts[0] == 961 and ts[-1] == 1 and len(ts) == 49 and ratio == 20
    assert s.timesteps(50, 1.0)[0][0] == 981
    st = default_adaptive_mask_settings(50)
    assert [st.dilate_scheduler(i) for i in (0, 5, 10, 15, 20, 25, 30, 35, 48)] == [20, 10, 5, 4, 3, 2, 1, 0, 0]
    assert sum(st.provoke_scheduler(i) for i in range(49)) == 21                 # 21 adapt calls in 49 steps
    assert st.provoke_scheduler(1) and not st.provoke_scheduler(0) and st.provoke_scheduler(44)


def test_adaptive_mask_loop_tiny_models(dev):
    from PIL import Image
    from coma_b200.inpaint.pipeline import (AdaptiveMaskInpaintPipeline, AdaptiveMaskSettings, DDIMSchedule, MaskDilateScheduler,
                                            ProvokeScheduler)
    from coma_b200.inpaint.segmenter import LuminanceSegmenter
    from coma_b200.inpaint.unet import UNet
    from coma_b200.inpaint.vae import VAE
    from oracle import inpaint_loop_oracle as lo
    from oracle import sd_oracle as so
    ucfg, vcfg = so.tiny_unet_cfg(), so.tiny_vae_cfg()
    usd = so.round_weights_fp16(so.make_unet_state_dict(0, ucfg))
    vsd = so.round_weights_fp16(so.make_vae_state_dict(1, vcfg))
    rng = np.random.default_rng(0)
    H = W = 64
    image = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    default = np.zeros((H, W), np.uint8)
    default[8:56, 16:48] = 255
    steps, strength, guidance, thres = 10, 0.9, 7.0, 0.0002
    settings = AdaptiveMaskSettings(MaskDilateScheduler(20, steps, [3, 3, 2, 2, 1, 1, 0, 0, 0, 0]), ProvokeScheduler(steps, [2, 4, 7], False))
    seg = LuminanceSegmenter(128)
    g = torch.Generator().manual_seed(5)
    pe = (torch.randn((77, ucfg["cross_attention_dim"]), generator=g) * 0.5).half()
    ne = (torch.randn((77, ucfg["cross_attention_dim"]), generator=g) * 0.5).half()

    pipe = AdaptiveMaskInpaintPipeline(UNet(usd, ucfg, dev), VAE(vsd, vcfg, dev))
    pipe.register_adaptive_mask_model(seg)
    pipe.register_adaptive_mask_settings(settings)
    B = 2
    gens = [torch.Generator(device=dev).manual_seed(100 + b) for b in range(B)]
    out = pipe(image=Image.fromarray(image), default_mask_image=Image.fromarray(default), prompt_embeds=pe, negative_prompt_embeds=ne,
               guidance_scale=guidance, strength=strength, num_inference_steps=steps, generator=gens, enforce_full_mask_ratio=0.0,
               human_detection_thres=thres, batch_size=B, return_trace=True, output_type="pt")
    assert out.images.shape == (B, H, W, 3) and torch.isfinite(out.images).all()

    ts, ratio = DDIMSchedule().timesteps(steps, strength)
    dv = lambda sd: {k: v.to(dev) for k, v in sd.items()}
    usd_d, vsd_d = dv(usd), dv(vsd)
    ctx2 = torch.stack([ne.float(), pe.float()]).to(dev)
    n_adapt = sum(settings.provoke_scheduler(i) for i in range(len(ts)))
    for b in range(B):
        g2 = torch.Generator(device=dev).manual_seed(100 + b)                  # the same stream the pipeline consumed
        draws = [torch.randn((1, 4, H // 8, W // 8), generator=g2, device=dev, dtype=torch.float16)[0].float() for _ in range(3 + n_adapt)]
        ref, final = lo.run_loop(usd_d, vsd_d, ucfg, vcfg, image, default, ctx2, ts, ratio, guidance, strength, draws, seg, settings,
                                 thres, 0.0, emulate_fp16=True, device=dev)
        h = H // 8
        for i, (mine, r) in enumerate(zip(out.trace, ref)):
            m_lat = mine["latents"].reshape(B, h, h, 4)[b].permute(2, 0, 1)
            scale = r["latents"].abs().max().item()
            err = (m_lat - r["latents"][0]).abs().max().item()
            # the adaptive mask is a hard decision on decoded pixels: require agreement of >= 98 % of the /8 mask cells,
            # and latents within 3 % of their scale (fp16 storage on both sides, error compounding over the steps)
            agree = (mine["mask64"].reshape(B, h, h)[b] == r["mask64"][0, 0]).float().mean().item()
            assert agree >= 0.98, (i, agree)
            assert err <= 3e-2 * scale, (i, err / scale)
        err = (out.images[b].permute(2, 0, 1) - final[0]).abs().max().item()
        assert err <= 5e-2, err


def test_visualization_dumps(dev, tmp_path):
    """`visualization_save_dir` + a segmenter with `use_visualizer = True` (utils/adaptive_mask_inpainting.py:1051-1060): one mask PNG
    (the reference's grey ramp) and one decoded-x0 PNG per provoke step, per batch element."""
    from PIL import Image
    from coma_b200.inpaint.pipeline import AdaptiveMaskInpaintPipeline, AdaptiveMaskSettings, MaskDilateScheduler, ProvokeScheduler
    from coma_b200.inpaint.segmenter import LuminanceSegmenter
    from coma_b200.inpaint.unet import UNet
    from coma_b200.inpaint.vae import VAE
    from oracle import sd_oracle as so
    ucfg, vcfg = so.tiny_unet_cfg(), so.tiny_vae_cfg()
    pipe = AdaptiveMaskInpaintPipeline(UNet(so.make_unet_state_dict(0, ucfg), ucfg, dev), VAE(so.make_vae_state_dict(1, vcfg), vcfg, dev))
    seg = LuminanceSegmenter(128)
    seg.use_visualizer = True
    pipe.register_adaptive_mask_model(seg)
    pipe.register_adaptive_mask_settings(AdaptiveMaskSettings(MaskDilateScheduler(20, 6, [2, 2, 1, 1, 0, 0]), ProvokeScheduler(6, [2, 4], False)))
    rng = np.random.default_rng(1)
    image = rng.integers(0, 256, (64, 64, 3), dtype=np.uint8)
    default = np.zeros((64, 64), np.uint8)
    default[8:56, 16:48] = 255
    g = torch.Generator().manual_seed(5)
    pe = (torch.randn((77, ucfg["cross_attention_dim"]), generator=g) * 0.5).half()
    out_dir = str(tmp_path / "vis")
    pipe(image=image, default_mask_image=default, prompt_embeds=pe, negative_prompt_embeds=torch.zeros_like(pe), guidance_scale=7.0, strength=1.0,
         num_inference_steps=6, generator=[torch.Generator(device=dev).manual_seed(b) for b in range(2)], enforce_full_mask_ratio=0.0,
         human_detection_thres=0.0002, batch_size=2, visualization_save_dir=out_dir, output_type="np")
    for root in (out_dir, os.path.join(out_dir, "1")):
        assert sorted(os.listdir(os.path.join(root, "masks"))) == ["00001.png", "00003.png"]      # provoke steps 2 and 4, 0-indexed loop counter
        assert sorted(os.listdir(os.path.join(root, "images"))) == ["00001.png", "00003.png"]
        m = np.asarray(Image.open(os.path.join(root, "masks", "00001.png")))
        assert m.shape == (64, 64) and set(np.unique(m)) <= {153, 255}                           # clip(0.6 + (1 - mask)) * 255
        assert np.asarray(Image.open(os.path.join(root, "images", "00003.png"))).shape == (64, 64, 3)
