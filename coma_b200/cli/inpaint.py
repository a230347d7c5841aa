"""2D HOI image synthesis driver — drop-in for the reference's `src/generation/inpaint.py` (work enumeration :187-269,
contiguous work-list slice :272-278, per-item loop :279-352, defaults :356-372, flags :376-409) on the B200 pipeline.

Same directory layout in and out:
  results/generation/asset_renders/<sc>/<c>/<asset>/<view>.png, asset_masks/.../<view>/<mask>.png (+ <view>.pickle with
  valid_mask_ids), asset_segs/..., prompts/<sc>/<c>/<asset>/prompts.pickle
  -> results/generation/inpaintings/<sc>/<c>/<asset>/<view>/<mask>/<prompt>/<id:06>.png
B200 extension: consecutive work items that share (render, mask, prompt, settings) and differ only by `inpaint_id` run as
ONE batched pipeline call (each item keeps its own seed = inpaint_id, its own adaptive mask and masked-image latents)."""
import os
import pickle
from glob import glob

import numpy as np
import torch
from PIL import Image

from constants.generation.prompts import ALLOWED_VIEWPOINT_AUGMENTATIONS, SC2DIFFUSERCONFIG, SCV2DIFFUSERCONFIG
from coma_b200 import dist as cdist


def prepare_asset_render_pths(asset_render_dir, supercategories, categories):
    """utils/prepare_renders.py:6-32 (without the hard-coded per-dataset asset allow-list)."""
    pths = sorted(glob(f"{asset_render_dir}/*/*/*/*.png"))
    if supercategories is not None:
        pths = [p for p in pths if p.split("/")[-4].lower() in supercategories]
    if categories is not None:
        pths = [p for p in pths if p.split("/")[-3].lower() in categories]
    return pths


def _cfg(supercategory, category, view_id, key, default):
    base = SC2DIFFUSERCONFIG.get(supercategory, {}).get(category, {})
    return SCV2DIFFUSERCONFIG.get(supercategory, {}).get(category, {}).get(view_id, base).get(key, base.get(key, default))


def enumerate_work(num_img_per_combination, supercategories, categories, asset_render_dir, asset_mask_dir, asset_seg_dir, prompts_dir,
                   save_dir, negative_prompt, defaults, debug=False):
    items = []
    for render in prepare_asset_render_pths(asset_render_dir, supercategories, categories):
        sc_s, c_s, asset_id, view_ext = render.split("/")[-4:]
        sc, c = sc_s.replace(":", "/"), c_s.replace(":", "/")
        view_id, ext = view_ext.split(".")
        assert ext == "png", "Rendering must have '.png' extension"
        meta = f"{asset_mask_dir}/{sc_s}/{c_s}/{asset_id}/{view_id}.pickle"
        if os.path.exists(meta):
            with open(meta, "rb") as fh:
                mask_ids = pickle.load(fh)["valid_mask_ids"]
        else:
            assert debug, "THIS SHOULD BE ONLY ALLOWED IN DEBUGGING MODE. RUN STEP2 PRIOR"
            mask_ids = [p.split("/")[-1].split(".")[0] for p in sorted(glob(f"{asset_mask_dir}/{sc_s}/{c_s}/{asset_id}/{view_id}/*"))]
        with open(f"{prompts_dir}/{sc}/{c}/{asset_id}/prompts.pickle", "rb") as fh:
            prompts = pickle.load(fh)["prompts"]
        augs = _cfg(sc, c, view_id, "view_text", ["original"])
        for mask_id in mask_ids:
            for prompt in prompts:
                for aug in augs:
                    assert aug in ALLOWED_VIEWPOINT_AUGMENTATIONS, f"viewpoint augmentation: '{aug}' not allowed"
                    if aug == "original":
                        text = prompt
                    elif aug != ", full body":
                        continue
                    else:
                        text = prompt + aug
                    out_dir = f"{save_dir}/{sc_s}/{c_s}/{asset_id}/{view_id}/{mask_id}/{text}"
                    for inpaint_id in range(num_img_per_combination):
                        items.append(dict(
                            asset_render_pth=render, asset_mask_pth=f"{asset_mask_dir}/{sc_s}/{c_s}/{asset_id}/{view_id}/{mask_id}.png",
                            asset_seg_pth=f"{asset_seg_dir}/{sc_s}/{c_s}/{asset_id}/{view_id}.png", result_save_dir=out_dir,
                            result_save_pth=f"{out_dir}/{inpaint_id:06}.png", inpaint_id=inpaint_id, input_prompt=text,
                            input_negprompt=negative_prompt,
                            ddim_steps=_cfg(sc, c, view_id, "ddim_steps", defaults["ddim_steps"]),
                            cfg_scale=_cfg(sc, c, view_id, "cfg_scale", defaults["cfg_scale"]),
                            strength=_cfg(sc, c, view_id, "strength", defaults["strength"]),
                            enforce_full_mask_ratio=_cfg(sc, c, view_id, "enforce_full_mask_ratio", defaults["enforce_full_mask_ratio"]),
                            human_detection_thres=_cfg(sc, c, view_id, "human_detection_thres", defaults["human_detection_thres"])))
    return sorted(items, key=lambda x: x["result_save_pth"])


def group_batches(items, max_batch):
    """Consecutive items (sorted by output path) that share everything but `inpaint_id`."""
    key = lambda it: tuple(it[k] for k in ("asset_render_pth", "asset_mask_pth", "input_prompt", "ddim_steps", "cfg_scale", "strength",
                                           "enforce_full_mask_ratio", "human_detection_thres"))
    out = []
    for it in items:
        if out and key(out[-1][0]) == key(it) and len(out[-1]) < max_batch:
            out[-1].append(it)
        else:
            out.append([it])
    return out


def inpaint_human(pipeline, embed_fn, num_img_per_combination, supercategories, categories, asset_render_dir, asset_mask_dir,
                  asset_seg_dir, prompts_dir, save_dir, negative_prompt, defaults, skip_done, verbose, parallel_num, parallel_idx,
                  batch_size=8, debug=False):
    """pipeline: coma_b200 AdaptiveMaskInpaintPipeline with the segmenter and settings registered (set_pipeline);
    embed_fn(text) -> [77, cross_dim] prompt embeddings (CLIP text encoder)."""
    items = enumerate_work(num_img_per_combination, supercategories, categories, asset_render_dir, asset_mask_dir, asset_seg_dir,
                           prompts_dir, save_dir, negative_prompt, defaults, debug)
    lo, hi = cdist.work_item_slice(len(items), parallel_idx, parallel_num)   # the reference's slice rule (:272-278)
    todo = []
    for it in items[lo:hi]:
        os.makedirs(it["result_save_dir"], exist_ok=True)
        if os.path.exists(it["result_save_pth"]) and skip_done:
            if verbose:
                print(f"Continueing {it['result_save_pth']} Since Already Done!")
            continue
        todo.append(it)
    done = 0
    # PNG encoding (zlib, ~20-40 ms per 512^2 image) runs on a few host threads behind the next batch's GPU work instead of between two
    # pipeline calls (the reference saves inline, src/generation/inpaint.py:350-352). A file appears under its final name
    # only when it is complete, so an interrupted run never leaves a truncated PNG for --skip_done to trust.
    writer = _PngWriter()
    for batch in group_batches(todo, batch_size):
        it = batch[0]
        init_image = Image.open(it["asset_render_pth"]).convert("RGB")
        default_mask_image = Image.open(it["asset_mask_pth"]).convert("L")
        gens = [torch.Generator(device=pipeline.dev).manual_seed(b["inpaint_id"]) for b in batch]   # :308-309
        res = pipeline(prompt_embeds=embed_fn(it["input_prompt"]), negative_prompt_embeds=embed_fn(it["input_negprompt"]),
                       image=init_image, default_mask_image=default_mask_image, guidance_scale=it["cfg_scale"], strength=it["strength"],
                       use_adaptive_mask=True, generator=gens, num_inference_steps=it["ddim_steps"],
                       enforce_full_mask_ratio=it["enforce_full_mask_ratio"], human_detection_thres=it["human_detection_thres"],
                       batch_size=len(batch))
        for b, img in zip(batch, res.images):
            writer.submit(img, b["result_save_pth"])
            done += 1
    writer.close()      # every file is on disk (or its error raised) before the caller sees the count
    return done


class _PngWriter:
    """Background PNG writer: `submit` returns at once, `close` waits for every file and re-raises the first failure."""

    def __init__(self, workers=4, max_pending=64):
        from concurrent.futures import ThreadPoolExecutor
        self._pool = ThreadPoolExecutor(max_workers=workers, thread_name_prefix="png")
        self._pending, self._max_pending = [], max_pending

    @staticmethod
    def _write(img, path):
        tmp = f"{path}.tmp{os.getpid()}"
        try:
            img.save(tmp, format="PNG")
            os.replace(tmp, path)
        except BaseException:
            if os.path.exists(tmp):
                os.remove(tmp)
            raise

    def submit(self, img, path):
        self._pending.append(self._pool.submit(self._write, img, path))
        if len(self._pending) >= self._max_pending:      # bound the images held in memory if the disk is slower than the GPU
            self._pending.pop(0).result()

    def close(self):
        try:
            for f in self._pending:
                f.result()
        finally:
            self._pending = []
            self._pool.shutdown(wait=True)


def load_state_dict(path):
    """diffusers-format weights: <dir>/diffusion_pytorch_model.safetensors (or .bin)."""
    st = os.path.join(path, "diffusion_pytorch_model.safetensors")
    if os.path.exists(st):
        from safetensors.torch import load_file
        return load_file(st)
    return torch.load(os.path.join(path, "diffusion_pytorch_model.bin"), map_location="cpu")


def read_diffusers_config(model_dir, sub):
    """<model_dir>/<sub>/config.json (diffusers layout) -> the cfg dict of coma_b200.inpaint.{unet.UNet, vae.VAE}, or None when the
    file is absent (the SD-1.5-inpainting / SD VAE defaults then apply). diffusers quirk kept: `attention_head_dim` of the SD-1.x UNet
    config is the NUMBER of heads (8)."""
    import json
    pth = os.path.join(model_dir, sub, "config.json")
    if not os.path.exists(pth):
        return None
    with open(pth) as fh:
        c = json.load(fh)
    if sub == "unet":
        down = c.get("down_block_types", ["CrossAttnDownBlock2D"] * 3 + ["DownBlock2D"])
        heads = c.get("attention_head_dim", 8)
        return dict(in_channels=c.get("in_channels", 9), out_channels=c.get("out_channels", 4),
                    block_out_channels=tuple(c.get("block_out_channels", (320, 640, 1280, 1280))), layers_per_block=c.get("layers_per_block", 2),
                    heads=heads[0] if isinstance(heads, (list, tuple)) else heads, cross_attention_dim=c.get("cross_attention_dim", 768),
                    groups=c.get("norm_num_groups", 32), attn_levels=tuple("CrossAttn" in t for t in down))
    return dict(in_channels=c.get("in_channels", 3), latent_channels=c.get("latent_channels", 4),
                block_out_channels=tuple(c.get("block_out_channels", (128, 256, 512, 512))), layers_per_block=c.get("layers_per_block", 2),
                groups=c.get("norm_num_groups", 32), scaling_factor=c.get("scaling_factor", 0.18215))


def build_segmenter(adaptive_mask_model_type, segmenter=None, default_pointrend_threshold=0.2):
    """The in-loop human segmenter (src/generation/inpaint.py:66-110) is a PLUG-IN here: detectron2 PointRend / SAM and their
    weights live outside this repository.
      --segmenter module:factory   import `module`, call `factory(adaptive_mask_model_type=..., pointrend_threshold=...)` and use
                                   the returned callable `(np.uint8 [H,W,3]) -> {"mask": np.uint8 [H,W], ...}` (the contract of
                                   PointRendPredictor.__call__, utils/adaptive_mask_inpainting.py:1225-1236);
      --adaptive_mask_model_type stub   the deterministic LuminanceSegmenter (benchmarks / smoke runs, clearly not a detector).
    Resolved BEFORE any weights are loaded, so a missing plug-in fails in milliseconds, not after the checkpoint is on the GPU."""
    if segmenter:
        import importlib
        mod, _, fn = segmenter.partition(":")
        if not mod or not fn:
            raise ValueError(f"--segmenter expects 'module:factory', got '{segmenter}'")
        factory = getattr(importlib.import_module(mod), fn)
        model = factory(adaptive_mask_model_type=adaptive_mask_model_type, pointrend_threshold=default_pointrend_threshold)
        if not callable(model):
            raise TypeError(f"{segmenter} returned {type(model).__name__}, expected a callable segmenter")
        return model
    if adaptive_mask_model_type == "stub":
        from coma_b200.inpaint.segmenter import LuminanceSegmenter
        return LuminanceSegmenter(128)
    raise NotImplementedError(
        f"--adaptive_mask_model_type {adaptive_mask_model_type}: the PointRend / SAM human segmenters of the reference need detectron2 / "
        "segment-anything and their weights, which are plug-ins outside this repository. Pass `--segmenter module:factory` (a factory "
        "returning a callable image -> {'mask': uint8 [H,W]}) or, for smoke runs, `--adaptive_mask_model_type stub`.")


def set_pipeline(model_dir, adaptive_mask_model_type, default_ddim_steps, default_pointrend_threshold=0.2, device="cuda", segmenter=None):
    """src/generation/inpaint.py:33-137: DDIM scheduler, fp16 inpainting checkpoint, segmenter, dilate / provoke schedules.
    `model_dir` is a local diffusers-layout directory (unet/, vae/, text_encoder/, tokenizer/); there is no network access."""
    from coma_b200.inpaint.pipeline import (AdaptiveMaskInpaintPipeline, AdaptiveMaskSettings, MaskDilateScheduler, ProvokeScheduler,
                                            default_adaptive_mask_settings)
    seg_model = build_segmenter(adaptive_mask_model_type, segmenter, default_pointrend_threshold)   # fails fast, before the weights
    from coma_b200.inpaint.unet import SD15_INPAINT, UNet
    from coma_b200.inpaint.vae import SD_VAE, VAE
    ucfg, vcfg = read_diffusers_config(model_dir, "unet") or SD15_INPAINT, read_diffusers_config(model_dir, "vae") or SD_VAE
    pipe = AdaptiveMaskInpaintPipeline(UNet(load_state_dict(os.path.join(model_dir, "unet")), ucfg, device=device),
                                       VAE(load_state_dict(os.path.join(model_dir, "vae")), vcfg, device=device))
    pipe.register_adaptive_mask_model(seg_model)
    n = int(default_ddim_steps * 0.1)
    if adaptive_mask_model_type in ("p", "stub"):
        settings = default_adaptive_mask_settings(default_ddim_steps)
    else:
        settings = AdaptiveMaskSettings(MaskDilateScheduler(20, default_ddim_steps, [10] * default_ddim_steps),
                                        ProvokeScheduler(default_ddim_steps, [] if adaptive_mask_model_type == "baseline" else
                                                         list(range(2, 11, 2)) + list(range(12, 41, 2)) + [45], False))
    pipe.register_adaptive_mask_settings(settings)
    return pipe


def clip_embedder(model_dir, device="cuda"):
    """CLIP-L/14 text encoder of the checkpoint (utils/adaptive_mask_inpainting.py:405-554, single prompt) on the B200 kernels
    (coma_b200/inpaint/clip.py); only the tokenizer (string processing) is transformers'."""
    from coma_b200.inpaint.clip import make_embedder
    return make_embedder(model_dir, device)
