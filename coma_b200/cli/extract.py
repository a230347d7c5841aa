"""ComA extraction / inference drivers — the drop-in CLI surface of `src/coma/extract_coma.py` (:66-503, flags :508-533)
and `src/coma/inference.py` (:26-147, flags :152-160) with the same directory layout and output files:

  results/coma/extracted_coma/<sc>/<c>/<asset>/<key>:<prompt>.{pickle,json}
  results/coma/affordance/<sc>/<c>/<asset>/<key>:<prompt>/{human_contact.npy,object_contact.ply,
                                                         orientational_tendency.npy,occupancy.npy}
  output/<sc>/<c>/{...}                                  (inference)

Sample ingest: the pickles of a SCAM are read by a thread pool, the vertex normals of ALL its samples are
one K6 launch and the downsample index gather runs on the GPU (`prepare_affordance_extraction_inputs_batch`); aggregation
and read-outs run through `utils.coma.ComA` / `utils.coma_occupancy.ComA_Occupancy` (coma_b200 kernels).
With torchrun (WORLD_SIZE > 1) every rank owns a block of HUMAN-VERTEX rows of the accumulators (`human_slice`), loads
1/world of the sample files and the staged fp32 samples are all-gathered, so no accumulator ever crosses GPUs; the
read-outs exchange per-vertex maps only, and `export` assembles the pickle on rank 0. Every rank takes part in the
(collective) read-outs; only rank 0 writes files.
"""
import json
import os
import pickle
from concurrent.futures import ThreadPoolExecutor
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


def _load_pickle(pth):
    with open(pth, "rb") as handle:
        return pickle.load(handle)


def prepare_affordance_extraction_inputs_batch(human_mesh_pths, human_downsample_metadata, object_downsample_metadata,
                                               human_use_downsample_pcd_raw, object_use_downsample_pcd_raw, eps,
                                               standardize_human_scale=False, scaler_range=None, camera_pth=None,
                                               human_params_pths=None, io_threads=8):
    """utils/coma.py:649-791 for a LIST of sample pickles sharing one topology (one SCAM): threaded pickle IO, ONE K6 launch for
    the vertex normals of every sample, the `downsample_indices` gather on the GPU. Returns a list (one entry per path, None for
    samples the scale filter drops) of the same dicts `prepare_affordance_extraction_inputs` returns — bit-identical values."""
    import torch
    assert not human_use_downsample_pcd_raw, "Human must use 'mesh' for Representation. You'll know why"
    if not human_mesh_pths:
        return []
    with ThreadPoolExecutor(max_workers=io_threads) as ex:
        datas = list(ex.map(_load_pickle, human_mesh_pths))
    keep = [True] * len(datas)
    if standardize_human_scale:
        cam_scale = _load_pickle(camera_pth)["scale"]
        for i, pp in enumerate(human_params_pths):
            hp_ = _load_pickle(pp)
            scaler = (512 / cam_scale) * (hp_["convert_data"]["z_mean"] / hp_["convert_data"]["focals"][0])
            keep[i] = scaler_range is None or (scaler_range[0] <= scaler <= scaler_range[1])
    obj_verts_orig = object_downsample_metadata["obj_vertices_original"]
    obj_vertex_normals_orig = normalize_vectors_np(np.asarray(object_downsample_metadata["obj_vertex_normals_original"]))
    hidx = np.asarray(human_downsample_metadata["downsample_indices"], dtype=np.int64)
    oidx = object_downsample_metadata["downsample_indices"]
    if object_use_downsample_pcd_raw:
        obj_verts = np.asarray(object_downsample_metadata["downsampled_pcd_points_raw"])
        obj_vertex_normals = np.asarray(object_downsample_metadata["downsampled_pcd_normal_raw"])
        assert len(obj_verts) == object_downsample_metadata["N_raw"]
    else:
        obj_verts = np.asarray(obj_verts_orig).copy()[oidx]
        obj_vertex_normals = obj_vertex_normals_orig.copy()[oidx]
        assert len(obj_verts) == object_downsample_metadata["N"]
    out = [None] * len(datas)
    kept = [i for i, k in enumerate(keep) if k]
    # group by topology (all SMPL-X samples share one face array; be robust to a stray different one)
    groups = {}
    for i in kept:
        f = np.asarray(datas[i]["faces"])
        groups.setdefault((f.shape, hash(np.ascontiguousarray(f).tobytes())), []).append(i)
    for members in groups.values():
        verts = np.stack([np.asarray(datas[i]["verts"], dtype=np.float64) for i in members])           # [S,V,3]
        mn = ingest.mesh_normals_for(datas[members[0]]["faces"], verts.shape[1])
        vt = torch.from_numpy(verts).to(mn.dev)
        normals = mn(vt, eps)                                                                          # one K6 launch, fp64
        it = torch.from_numpy(hidx).to(mn.dev)
        hv = vt.index_select(1, it).cpu().numpy()
        hn = normals.index_select(1, it).cpu().numpy()
        assert hv.shape[1] == human_downsample_metadata["N"]
        for j, i in enumerate(members):
            out[i] = dict(human_verts=hv[j], human_vertex_normals=hn[j], obj_verts=obj_verts, obj_vertex_normals=obj_vertex_normals)
    return out


def _build_coma(visualize_type, H, O, hp, scale_tolerance, device="cuda", human_slice=None):
    from utils.coma import ComA
    from utils.coma_occupancy import ComA_Occupancy
    common = dict(human_res=H, obj_res=O, normal_res=hp["normal_res"], spatial_res=hp["spatial_res"],
                  proximity_settings=dict(spatial_grid_size=hp["spatial_grid_size"], spatial_grid_thres=hp["spatial_grid_thres"]),
                  principle_vec=hp["principle_vec"], sub_principle_vec=hp["sub_principle_vec"],
                  rel_dist_method=hp["rel_dist_method"], normal_gaussian_sigma=hp["normal_gaussian_sigma"], eps=hp["eps"],
                  device=device, human_slice=human_slice)
    if visualize_type == "occupancy":
        return ComA_Occupancy(scale_tolerance=scale_tolerance, **common)
    return ComA(**common)


def write_affordance(coma, visualize_type, hp, out_dir, object_downsample_metadata, save=True):
    """The four read-outs of src/coma/extract_coma.py:428-483 == src/coma/inference.py:95-147.
    On a row-sharded instance the read-outs are COLLECTIVE: every rank must call this; only `save=True` ranks write."""
    from utils.coma import get_aggregated_contact
    if save:
        os.makedirs(out_dir, exist_ok=True)
    if visualize_type == "aggr-human-contact":
        agg, _ = get_aggregated_contact(coma=coma, contact_map_type="human", significant_contact_ratio=hp["significant_contact_ratio"])
        if save:
            np.save(f"{out_dir}/human_contact.npy", agg / agg.max())
    elif visualize_type == "aggr-object-contact":
        agg, _ = get_aggregated_contact(coma=coma, contact_map_type="obj", significant_contact_ratio=hp["significant_contact_ratio"])
        score = agg / agg.max()
        if save:
            write_point_cloud_ply(f"{out_dir}/object_contact.ply", object_downsample_metadata["downsampled_pcd_points_raw"],
                                  object_downsample_metadata["downsampled_pcd_normal_raw"], jet_rgb(score))
    elif visualize_type == "orientation":
        s = coma.compute_nonphysical_response_sphere(n_bin=1e6, nonphysical_type="human", as_numpy=True)["human"][:, 0]
        if save:
            np.save(f"{out_dir}/orientational_tendency.npy", (s - s.min()) / (s.max() - s.min()))
    elif visualize_type == "occupancy":
        prob_field = coma.return_aggregated_spatial_grids(human_indices=None).cpu().numpy()
        prob_field /= prob_field.max()
        prob_field = 0.7 * prob_field
        if save:
            np.save(f"{out_dir}/occupancy.npy", dict(prob_field=prob_field, spatial_grid_metadata=coma.spatial_grid_metadata))


def run_affordance_extraction(supercategories, categories, prompts, camera_dir, human_params_dir, asset_downsample_dir,
                              human_postfilter_dir, human_sample_dir, coma_save_dir, affordance_save_dir, hyperparams,
                              hyperparams_key, scale_tolerance=3.0, skip_done=False,
                              smplx_downsample_dir="./constants/mesh", device="cuda", **_unused):
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

        # multi-GPU: this rank owns a block of human-vertex rows of every accumulator and loads 1/world of the sample files
        coma = _build_coma(visualize_type, H, O, hp, scale_tolerance, device=device,
                           human_slice=cdist.human_slice(H, rank, world) if world > 1 else None)
        done = skip_done and os.path.exists(save_pth)
        if done:
            coma.load(save_pth)  # the ComA pickle is the checkpoint (:350-351); a sharded rank keeps its rows only
        else:
            mine = [inputs[i] for i in cdist.sample_shard(len(inputs), rank, world)]
            hparams = []
            for pth in mine:
                sc_s2, c_s2, asset2, view_id, mask_id, prompt, id_ext = pth.split("/")[-7:]
                hparams.append(f"{human_params_dir}/{sc_s2}/{c_s2}/{asset2}/{view_id}/{mask_id}/{prompt.replace('total:', '')}/{id_ext}")
            xs = prepare_affordance_extraction_inputs_batch(mine, human_md, object_md, hp["human_use_downsample_pcd_raw"],
                                                            hp["object_use_downsample_pcd_raw"], hp["eps"], hp["standardize_human_scale"],
                                                            hp["scaler_range"], camera_pth, hparams)
            for x in xs:
                if x is None:
                    continue
                coma.register_sample_to_cache(human_verts=x["human_verts"], human_normals=x["human_vertex_normals"],
                                              obj_verts=x["obj_verts"], obj_normals=x["obj_vertex_normals"])
            coma.aggregate_all_samples(exchange=world > 1)
            if rank == 0:
                os.makedirs(save_dir, exist_ok=True)
            coma.export(save_pth=save_pth)       # collective on a sharded instance; rank 0 writes
        write_affordance(coma, visualize_type, hp, f"{affordance_save_dir}/{sc}/{c}/{asset}/{hyperparams_key}:{main}", object_md,
                         save=rank == 0)
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
