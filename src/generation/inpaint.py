"""CLI of the 2D HOI image synthesis stage — same flags and defaults as the reference's src/generation/inpaint.py:356-453,
plus --model_dir (local diffusers-layout checkpoint; no network here) and --batch_size (B200 batching of seeds)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from constants.generation.inpaint_ldm import HF_MODEL_KEYS  # noqa: E402
from constants.metadata import DEFAULT_SEED  # noqa: E402

NEGATIVE_PROMPT = "worst quality, normal quality, low quality, bad anatomy, artifacts, blurry, cropped, watermark, greyscale, nsfw"

if __name__ == "__main__":
    p = argparse.ArgumentParser()
    p.add_argument("--num_img_per_combination", type=int, default=10)
    p.add_argument("--supercategories", type=str, nargs="+")
    p.add_argument("--categories", type=str, nargs="+")
    p.add_argument("--asset_render_dir", type=str, default="results/generation/asset_renders")
    p.add_argument("--asset_mask_dir", type=str, default="results/generation/asset_masks")
    p.add_argument("--asset_seg_dir", type=str, default="results/generation/asset_segs")
    p.add_argument("--prompts_dir", type=str, default="results/generation/prompts")
    p.add_argument("--save_dir", type=str, default="results/generation/inpaintings")
    p.add_argument("--ldm_model_key", type=str, default="realisticvision", choices=HF_MODEL_KEYS.keys())
    p.add_argument("--model_dir", type=str, default=None, help="local diffusers-layout checkpoint of --ldm_model_key")
    p.add_argument("--adaptive_mask_model_type", type=str, choices=["baseline", "p", "ps", "ps_ae", "s_pdb_ae", "s_db_ae", "s_ab_ae", "stub"], default="p")
    p.add_argument("--default_cfg_scale", type=float, default=11.0)
    p.add_argument("--default_strength", type=float, default=0.98)
    p.add_argument("--default_ddim_steps", type=int, default=50)
    p.add_argument("--default_pointrend_threshold", type=float, default=0.2)
    p.add_argument("--default_enforce_full_mask_ratio", type=float, default=0.0)
    p.add_argument("--default_human_detection_thres", type=float, default=0.015)
    p.add_argument("--enable_sam_multitask_output", action="store_true")
    p.add_argument("--negative_prompt", type=str, default=NEGATIVE_PROMPT)
    p.add_argument("--enable_safety_checker", action="store_true")
    p.add_argument("--use_visualizer", action="store_true")
    p.add_argument("--skip_done", action="store_true", default=True)
    p.add_argument("--no_skip_done", dest="skip_done", action="store_false", help="regenerate images that already exist")
    p.add_argument("--segmenter", type=str, default=None,
                   help="plug-in human segmenter 'module:factory' (see coma_b200.cli.inpaint.build_segmenter); required unless "
                        "--adaptive_mask_model_type stub")
    p.add_argument("--verbose", action="store_true")
    p.add_argument("--seed", type=int, default=DEFAULT_SEED)
    p.add_argument("--parallel_num", type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
    p.add_argument("--parallel_idx", type=int, default=int(os.environ.get("RANK", "0")))
    p.add_argument("--batch_size", type=int, default=8)
    a = p.parse_args()
    low = lambda xs: None if xs is None else [x.lower() for x in xs]
    if a.adaptive_mask_model_type == "baseline":
        a.save_dir = f"{a.save_dir}_noadaptivemask"
    from coma_b200.cli.inpaint import clip_embedder, inpaint_human, set_pipeline
    from coma_b200.cli.io import seed_everything
    import torch
    seed_everything(a.seed)
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    model_dir = a.model_dir or os.path.join("checkpoints", HF_MODEL_KEYS[a.ldm_model_key].replace("/", "--"))
    pipe = set_pipeline(model_dir, a.adaptive_mask_model_type, a.default_ddim_steps, a.default_pointrend_threshold, segmenter=a.segmenter)
    inpaint_human(pipe, clip_embedder(model_dir), a.num_img_per_combination, low(a.supercategories), low(a.categories), a.asset_render_dir,
                  a.asset_mask_dir, a.asset_seg_dir, a.prompts_dir, a.save_dir, a.negative_prompt,
                  dict(ddim_steps=a.default_ddim_steps, cfg_scale=a.default_cfg_scale, strength=a.default_strength,
                       enforce_full_mask_ratio=a.default_enforce_full_mask_ratio, human_detection_thres=a.default_human_detection_thres),
                  a.skip_done, a.verbose, a.parallel_num, a.parallel_idx, a.batch_size)
