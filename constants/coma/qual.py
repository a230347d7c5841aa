"""Hyper-parameter presets of the qualitative ComA extraction (values identical to the reference's
constants/coma/qual.py:1-75; every preset is the base preset overridden by its own entries)."""

_BASE = dict(
    human_res="FULL", human_use_downsample_pcd_raw=False, object_res="180", object_use_downsample_pcd_raw=True,
    principle_vec=[0, 0, 1], sub_principle_vec=[0, 1, 0], rel_dist_method="dist",
    spatial_grid_size=0.06, spatial_grid_thres=0.24, normal_gaussian_sigma=0.2, normal_res=250, spatial_res=0,
    eps=1e-10, significant_contact_ratio=0.3, enable_postfilter=True, standardize_human_scale=False,
    scaler_range=(0.75, 1.25), visualize_type="aggr-human-contact", vis_example_num=0, quant_mode=False, quant_keys=[],
)

_OVERRIDES = {
    "qual:001": dict(),
    "qual:backpack_human_contact": dict(spatial_grid_size=0.07, spatial_grid_thres=0.03, normal_gaussian_sigma=0.25,
                                        significant_contact_ratio=0.1, visualize_type="aggr-human-contact"),
    "qual:backpack_object_contact": dict(spatial_grid_size=0.15, spatial_grid_thres=0.05, normal_gaussian_sigma=0.25,
                                         significant_contact_ratio=0.1, human_res="1000", object_res="1500",
                                         visualize_type="aggr-object-contact"),
    "qual:backpack_occupancy": dict(spatial_res=30, normal_res=0, human_res="FULL", object_res="1500",
                                    object_use_downsample_pcd_raw=False, visualize_type="occupancy"),
    "qual:backpack_orientation": dict(spatial_grid_size=0.03, spatial_grid_thres=0.1, normal_gaussian_sigma=0.2,
                                      significant_contact_ratio=0.1, visualize_type="orientation", vis_example_num=1),
}

QUAL_AFFORDANCE_EXTRACTION_HYPERPARAMS_DICT = {k: {**_BASE, **v} for k, v in _OVERRIDES.items()}

# scripts/learn_coma.sh asks for "qual:<category>_object" / "_human", names the reference's table does not contain
# (SURVEY Appendix D); accept them as aliases of the "_object_contact" / "_human_contact" presets.
for _k in list(QUAL_AFFORDANCE_EXTRACTION_HYPERPARAMS_DICT):
    if _k.endswith("_object_contact") or _k.endswith("_human_contact"):
        QUAL_AFFORDANCE_EXTRACTION_HYPERPARAMS_DICT.setdefault(_k[: -len("_contact")], QUAL_AFFORDANCE_EXTRACTION_HYPERPARAMS_DICT[_k])
