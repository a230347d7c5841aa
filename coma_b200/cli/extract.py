"""ComA extraction / inference drivers — the drop-in CLI surface of `src/coma/extract_coma.py` (:66-503, flags :508-533)
and `src/coma/inference.py` (:26-147, flags :152-160) with the same directory layout and output files:

  results/coma/extracted_coma/<sc>/<c>/<asset>/<key>:<prompt>.{pickle,json}
  results/coma/affordance/<sc>/<c>/<asset>/<key>:<prompt>/{human_contact.npy,object_contact.ply,
                                                         orientational_tendency.npy,occupancy.npy}
  output/<sc>/<c>/{...}                                  (inference)

The per-sample host work (pickle IO, vertex normals, index gather) stays on the CPU as in the reference; aggregation and
read-outs run on the GPU through `utils.coma.ComA` / `utils.coma_occupancy.ComA_Occupancy` (coma_b200 kernels).
With torchrun (WORLD_SIZE > 1) the samples of each SCAM are sharded over the ranks and the accumulators are summed with
one all-reduce before rank 0 writes the outputs.
"""
import json
import os
import pickle
from copy import deepcopy
from glob import glob

import numpy as np

from coma_b200 import dist as cdist
from coma_b200.cli.io import jet_rgb, vertex_normals, write_point_cloud_ply
from coma_b200 import ingest
from coma_b200.misc import normalize_vectors_np

ERROR_SENTINELS = ["NOT ALLOWED VIEWPOINT PROMPTS", "ERRONEOUS SAMPLE DUE TO TOO SMALL HUMAN", "TOO LITTLE INLIERS",
                   "LARGELY PENETRATED HUMAN"]  # src/coma/extract_coma.py:233-238


def prepare_affordance_extraction_inputs(human_mesh_pth, human_downsample_metadata, object_downsample_metadata,
                                         human_use_downsample_pcd_raw, object_use_downsample_pcd_raw, eps,
                                         standardize_human_scale=False, scaler_range=None, camera_pth=None,
                                         human_params_pth=None):
    """utils/coma.py:649-791 for `human_mesh_pth_type == "pickle"` (the only type the scripts use)."""
    with open(human_mesh_pth, "rb") as handle:
        human_data = pickle.load(handle)
    human_verts_orig, human_faces_orig = human_data["verts"], human_data["faces"]
    # K6: area-weighted vertex normals + normalize_vectors_np(., eps) on the GPU (fixed SMPL-X topology, cached corner list)
    human_vertex_normals_orig = ingest.vertex_normals(human_verts_orig, human_faces_orig, eps=eps)
    obj_verts_orig = object_downsample_metadata["obj_vertices_original"]
    obj_vertex_normals_orig = normalize_vectors_np(np.asarray(object_downsample_metadata["obj_vertex_normals_original"]))
    hidx = human_downsample_metadata["downsample_indices"]
    oidx = object_downsample_metadata["downsample_indices"]
    assert not human_use_downsample_pcd_raw, "Human must use 'mesh' for Representation. You'll know why"
    human_verts = np.asarray(human_verts_orig).copy()[hidx]
    human_vertex_normals = human_vertex_normals_orig.copy()[hidx]
    assert len(human_verts) == human_downsample_metadata["N"]
    if object_use_downsample_pcd_raw:
        obj_verts = object_downsample_metadata["downsampled_pcd_points_raw"]
        obj_vertex_normals = object_downsample_metadata["downsampled_pcd_normal_raw"]
        assert len(obj_verts) == object_downsample_metadata["N_raw"]
    else:
        obj_verts = np.asarray(obj_verts_orig).copy()[oidx]
        obj_vertex_normals = obj_vertex_normals_orig.copy()[oidx]
        assert len(obj_verts) == object_downsample_metadata["N"]
    if standardize_human_scale:
        with open(camera_pth, "rb") as handle:
            cam_scale = pickle.load(handle)["scale"]
        with open(human_params_pth, "rb") as handle:
            hp = pickle.load(handle)
        scaler = (512 / cam_scale) * (hp["convert_data"]["z_mean"] / hp["convert_data"]["focals"][0])
        if scaler_range is not None and (scaler < scaler_range[0] or scaler > scaler_range[1]):
            return None
    return dict(human_verts=np.asarray(human_verts), human_vertex_normals=np.asarray(human_vertex_normals),
                obj_verts=np.asarray(obj_verts), obj_vertex_normals=np.asarray(obj_vertex_normals))


def _build_coma(visualize_type, H, O, hp, scale_tolerance, device="cuda"):
    from utils.coma import ComA
    from utils.coma_occupancy import ComA_Occupancy
    common = dict(human_res=H, obj_res=O, normal_res=hp["normal_res"], spatial_res=hp["spatial_res"],
                  proximity_settings=dict(spatial_grid_size=hp["spatial_grid_size"], spatial_grid_thres=hp["spatial_grid_thres"]),
                  principle_vec=hp["principle_vec"], sub_principle_vec=hp["sub_principle_vec"],
                  rel_dist_method=hp["rel_dist_method"], normal_gaussian_sigma=hp["normal_gaussian_sigma"], eps=hp["eps"],
                  device=device)
    if visualize_type == "occupancy":
        return ComA_Occupancy(scale_tolerance=scale_tolerance, **common)
    return ComA(**common)


def write_affordance(coma, visualize_type, hp, out_dir, object_downsample_metadata):
    """The four read-outs of src/coma/extract_coma.py:428-483 == src/coma/inference.py:95-147."""
    from utils.coma import get_aggregated_contact
    os.makedirs(out_dir, exist_ok=True)
    if visualize_type == "aggr-human-contact":
        agg, _ = get_aggregated_contact(coma=coma, contact_map_type="human", significant_contact_ratio=hp["significant_contact_ratio"])
        np.save(f"{out_dir}/human_contact.npy", agg / agg.max())
    elif visualize_type == "aggr-object-contact":
        agg, _ = get_aggregated_contact(coma=coma, contact_map_type="obj", significant_contact_ratio=hp["significant_contact_ratio"])
        score = agg / agg.max()
        write_point_cloud_ply(f"{out_dir}/object_contact.ply", object_downsample_metadata["downsampled_pcd_points_raw"],
                              object_downsample_metadata["downsampled_pcd_normal_raw"], jet_rgb(score))
    elif visualize_type == "orientation":
        s = coma.compute_nonphysical_response_sphere(n_bin=1e6, nonphysical_type="human", as_numpy=True)["human"][:, 0]
        np.save(f"{out_dir}/orientational_tendency.npy", (s - s.min()) / (s.max() - s.min()))
    elif visualize_type == "occupancy":
        prob_field = coma.return_aggregated_spatial_grids(human_indices=None).cpu().numpy()
        prob_field /= prob_field.max()
        prob_field = 0.7 * prob_field
        np.save(f"{out_dir}/occupancy.npy", dict(prob_field=prob_field, spatial_grid_metadata=coma.spatial_grid_metadata))


def run_affordance_extraction(supercategories, categories, prompts, camera_dir, human_params_dir, asset_downsample_dir,
                              human_postfilter_dir, human_sample_dir, coma_save_dir, affordance_save_dir, hyperparams,
                              hyperparams_key, scale_tolerance=3.0, skip_done=False,
                              smplx_downsample_dir="./constants/mesh", **_unused):
    hp = hyperparams
    visualize_type, quant_mode = hp["visualize_type"], hp["quant_mode"]
    rank, world, _ = cdist.init_process_group()
    with open(f"{smplx_downsample_dir}/smplx_star_downsampled_{hp['human_res']}.pickle", "rb") as handle:
        human_md = pickle.load(handle)

    # ---- SCAMs = (supercategory, category, asset, main prompt), src/coma/extract_coma.py:147-173
    scams = set()
    for pth in sorted(set(glob(f"{human_sample_dir}/*/*/*/*/*/*/*.pickle"))):
        sc_s, c_s, asset, _, _, prompt, _ = pth.split("/")[-7:]
        sc, c = sc_s.replace(":", "/"), c_s.replace(":", "/")
        main = "total" if "total:" in prompt.split(",")[0] else prompt.split(",")[0]
        if (supercategories is not None and sc.lower() not in supercategories) or \
           (categories is not None and c.lower() not in categories) or (prompts is not None and main.lower() not in prompts):
            continue
        scams.add((sc, c, asset, main))

    remains = {}
    for sc, c, asset, main in sorted(scams):
        if quant_mode and main != "total":
            continue
        sc_s, c_s = sc.replace("/", ":"), c.replace("/", ":")
        inputs, camera_pth = [], None
        for pth in sorted(set(glob(f"{human_sample_dir}/{sc_s}/{c_s}/{asset}/*/*/{main}*/*.pickle"))):
            _, _, _, view_id, mask_id, prompt, id_ext = pth.split("/")[-7:]
            inpaint_id, ext = id_ext.split(".")
            assert ext == "pickle", "Human Finals must have '.pickle' extension"
            if hp["enable_postfilter"]:  # :29-63
                key = (sc, c, asset, main)
                if key not in remains:
                    fp = f"{human_postfilter_dir}/{sc_s}/{c_s}/{asset}/{main}.json"
                    assert os.path.exists(fp), fp
                    with open(fp) as rf:
                        remains[key] = {tuple(x) for x in json.load(rf)}
                if (view_id, mask_id, prompt, inpaint_id) not in remains[key]:
                    continue
            with open(pth, "rb") as handle:
                sample = pickle.load(handle)
            if isinstance(sample, str):  # sentinel-string error protocol
                assert sample in ERROR_SENTINELS, "What more errors could there be?"
                assert not hp["enable_postfilter"], pth
                continue
            camera_pth = f"{camera_dir}/{sc_s}/{c_s}/{asset}/{view_id}.pickle"
            inputs.append(pth)
        if not inputs:
            continue

        save_dir = f"{coma_save_dir}/{sc_s}/{c_s}/{asset}"
        json_pth, save_pth = f"{save_dir}/{hyperparams_key}:{main}.json", f"{save_dir}/{hyperparams_key}:{main}.pickle"
        with open(f"{asset_downsample_dir}/{sc_s}/{c_s}/{asset}_{hp['object_res']}.pickle", "rb") as handle:
            object_md = deepcopy(pickle.load(handle))
        H = human_md["N_raw"] if hp["human_use_downsample_pcd_raw"] else human_md["N"]
        O = object_md["N_raw"] if hp["object_use_downsample_pcd_raw"] else object_md["N"]
        if rank == 0 and not os.path.exists(json_pth):
            os.makedirs(save_dir, exist_ok=True)
            info = dict(input_human_pths=inputs, camera_pth=camera_pth, result_save_dir=save_dir, result_json_pth=json_pth,
                        result_save_pth=save_pth, H=H, O=O)
            info.update(hp)
            with open(json_pth, "w") as wf:
                json.dump(info, wf, indent=1)

        coma = _build_coma(visualize_type, H, O, hp, scale_tolerance)
        if skip_done and os.path.exists(save_pth):
            coma.load(save_pth)  # the ComA pickle is the checkpoint (:350-351)
        else:
            occupancy = visualize_type == "occupancy"
            mine = inputs if (occupancy or world == 1) else [inputs[i] for i in cdist.sample_shard(len(inputs), rank, world)]
            for pth in mine:
                sc_s2, c_s2, asset2, view_id, mask_id, prompt, id_ext = pth.split("/")[-7:]
                hparams = f"{human_params_dir}/{sc_s2}/{c_s2}/{asset2}/{view_id}/{mask_id}/{prompt.replace('total:', '')}/{id_ext}"
                x = prepare_affordance_extraction_inputs(pth, human_md, object_md, hp["human_use_downsample_pcd_raw"],
                                                         hp["object_use_downsample_pcd_raw"], hp["eps"], hp["standardize_human_scale"],
                                                         hp["scaler_range"], camera_pth, hparams)
                if x is None:
                    continue
                coma.register_sample_to_cache(human_verts=x["human_verts"], human_normals=x["human_vertex_normals"],
                                              obj_verts=x["obj_verts"], obj_normals=x["obj_vertex_normals"])
            coma.aggregate_all_samples()
            if not occupancy:
                coma.all_reduce()
            if rank == 0:
                os.makedirs(save_dir, exist_ok=True)
                coma.export(save_pth=save_pth)
        if rank == 0:
            write_affordance(coma, visualize_type, hp, f"{affordance_save_dir}/{sc}/{c}/{asset}/{hyperparams_key}:{main}", object_md)
        del coma


def inference(supercategory, category, coma_path, smplx_downsample_pth, asset_downsample_pth, visualize_type, hyperparams_key,
              hyperparams, output_dir):
    """src/coma/inference.py:26-147. NB the reference overrides the CLI's visualize_type with the preset's (:46)."""
    hp = hyperparams
    visualize_type = hp["visualize_type"]
    with open(smplx_downsample_pth, "rb") as handle:
        human_md = pickle.load(handle)
    with open(asset_downsample_pth, "rb") as handle:
        object_md = deepcopy(pickle.load(handle))
    H = human_md["N_raw"] if hp["human_use_downsample_pcd_raw"] else human_md["N"]
    O = object_md["N_raw"] if hp["object_use_downsample_pcd_raw"] else object_md["N"]
    coma = _build_coma(visualize_type, H, O, hp, scale_tolerance=3.0)
    coma.load(coma_path)
    write_affordance(coma, visualize_type, hp, f"{output_dir}/{supercategory}/{category}", object_md)
