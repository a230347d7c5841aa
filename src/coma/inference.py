"""CLI of the ComA inference stage — same flags as the reference's src/coma/inference.py:150-182 (whose import of the
non-existent `constants.coma.coma_basic_settings` makes it un-importable as shipped; non-"qual:" keys use the quant table)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from constants.coma.qual import QUAL_AFFORDANCE_EXTRACTION_HYPERPARAMS_DICT  # noqa: E402
from constants.coma.quant import QUANT_AFFORDANCE_EXTRACTION_HYPERPARAMS_DICT  # noqa: E402
from constants.metadata import DEFAULT_SEED  # noqa: E402

if __name__ == "__main__":
    p = argparse.ArgumentParser()
    p.add_argument("--supercategory", type=str)
    p.add_argument("--category", type=str)
    p.add_argument("--coma_path", type=str)
    p.add_argument("--visualize_type", type=str, choices=["aggr-human-contact", "aggr-object-contact", "orientation", "occupancy"])
    p.add_argument("--smplx_downsample_pth", type=str)
    p.add_argument("--asset_downsample_pth", type=str)
    p.add_argument("--hyperparams_key", type=str)
    p.add_argument("--output_dir", type=str, default="output")
    p.add_argument("--seed", type=int, default=DEFAULT_SEED)
    a = p.parse_args()
    from coma_b200.cli.extract import inference
    from coma_b200.cli.io import seed_everything
    seed_everything(a.seed)
    assert a.hyperparams_key is not None, "You must Specify the 'args.hypeparams_key'"
    table = QUAL_AFFORDANCE_EXTRACTION_HYPERPARAMS_DICT if "qual:" in a.hyperparams_key else QUANT_AFFORDANCE_EXTRACTION_HYPERPARAMS_DICT
    inference(supercategory=a.supercategory, category=a.category, coma_path=a.coma_path, visualize_type=a.visualize_type,
              smplx_downsample_pth=a.smplx_downsample_pth, asset_downsample_pth=a.asset_downsample_pth,
              hyperparams_key=a.hyperparams_key, hyperparams=table[a.hyperparams_key], output_dir=a.output_dir)
