"""CLI of the ComA extraction stage — same flags and defaults as the reference's src/coma/extract_coma.py:506-578."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from constants.coma.qual import QUAL_AFFORDANCE_EXTRACTION_HYPERPARAMS_DICT  # noqa: E402
from constants.coma.quant import QUANT_AFFORDANCE_EXTRACTION_HYPERPARAMS_DICT  # noqa: E402
from constants.metadata import DEFAULT_SEED  # noqa: E402

if __name__ == "__main__":
    p = argparse.ArgumentParser()
    p.add_argument("--supercategories", type=str, nargs="+")
    p.add_argument("--categories", type=str, nargs="+")
    p.add_argument("--prompts", type=str, nargs="+")
    p.add_argument("--camera_dir", type=str, default="results/generation/cameras")
    p.add_argument("--human_params_dir", type=str, default="results/generation/human_preds")
    p.add_argument("--asset_downsample_dir", type=str, default="results/coma/asset_downsample")
    p.add_argument("--human_postfilter_dir", type=str, default="results/coma/human_postfilterings")
    p.add_argument("--human_sample_dir", type=str, default="results/generation/human_sample")
    p.add_argument("--coma_save_dir", type=str, default="results/coma/extracted_coma")
    p.add_argument("--affordance_save_dir", type=str, default="results/coma/affordance")
    p.add_argument("--smplx_canon_obj_pth", type=str, default="./constants/mesh/smplx_star.obj")
    p.add_argument("--hyperparams_key", choices=list(QUANT_AFFORDANCE_EXTRACTION_HYPERPARAMS_DICT) + list(QUAL_AFFORDANCE_EXTRACTION_HYPERPARAMS_DICT))
    p.add_argument("--visualize", action="store_true")
    p.add_argument("--vis_example_num", type=int)
    p.add_argument("--interactive", action="store_true")
    p.add_argument("--vis_interactive", action="store_true")
    p.add_argument("--fovy", type=float, default=27.5)
    p.add_argument("--tmp_cache_dir", type=str, default="results/coma_tmp_cache")
    p.add_argument("--selected_object_indices", type=str, help="Type as '21 22' or '21-25'", default="")
    p.add_argument("--scale_tolerance", type=float, default=3.0)
    p.add_argument("--skip_done", action="store_true")
    p.add_argument("--seed", type=int, default=DEFAULT_SEED)
    a = p.parse_args()
    low = lambda xs: None if xs is None else [x.lower() for x in xs]
    from coma_b200.cli.extract import run_affordance_extraction
    from coma_b200.cli.io import seed_everything
    seed_everything(a.seed)
    assert a.hyperparams_key is not None, "You must Specify the 'args.hypeparams_key'"
    table = QUAL_AFFORDANCE_EXTRACTION_HYPERPARAMS_DICT if "qual:" in a.hyperparams_key else QUANT_AFFORDANCE_EXTRACTION_HYPERPARAMS_DICT
    run_affordance_extraction(
        supercategories=low(a.supercategories), categories=low(a.categories), prompts=low(a.prompts), camera_dir=a.camera_dir,
        human_params_dir=a.human_params_dir, asset_downsample_dir=a.asset_downsample_dir, human_postfilter_dir=a.human_postfilter_dir,
        human_sample_dir=a.human_sample_dir, coma_save_dir=a.coma_save_dir, affordance_save_dir=a.affordance_save_dir,
        hyperparams=table[a.hyperparams_key], hyperparams_key=a.hyperparams_key, scale_tolerance=a.scale_tolerance,
        skip_done=a.skip_done, smplx_downsample_dir=os.path.dirname(a.smplx_canon_obj_pth) or ".")
