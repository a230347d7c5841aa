"""CLI surface: presets, IO helpers (CPU) and an end-to-end extract -> inference run on a tiny synthetic tree (GPU)."""
import importlib.util
import os
import pickle

import numpy as np
import pytest


def test_presets_match_reference_values():
    from constants.coma.qual import QUAL_AFFORDANCE_EXTRACTION_HYPERPARAMS_DICT as Q
    from constants.coma.quant import QUANT_AFFORDANCE_EXTRACTION_HYPERPARAMS_DICT as QT
    p = Q["qual:backpack_human_contact"]
    assert (p["spatial_grid_size"], p["spatial_grid_thres"], p["normal_gaussian_sigma"], p["significant_contact_ratio"]) == (0.07, 0.03, 0.25, 0.1)
    assert p["human_res"] == "FULL" and p["object_res"] == "180" and p["normal_res"] == 250 and p["eps"] == 1e-10
    o = Q["qual:backpack_occupancy"]
    assert o["spatial_res"] == 30 and o["normal_res"] == 0 and o["visualize_type"] == "occupancy" and not o["object_use_downsample_pcd_raw"]
    assert Q["qual:backpack_object"] is Q["qual:backpack_object_contact"]      # learn_coma.sh spelling
    assert QT["quant:full"]["object_res"] == "2048" and QT["quant:full"]["quant_mode"]
    ref = "/root/reference/constants/coma/qual.py"
    if os.path.exists(ref):                                                    # dev container only
        spec = importlib.util.spec_from_file_location("refq", ref)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        for k, v in m.QUAL_AFFORDANCE_EXTRACTION_HYPERPARAMS_DICT.items():
            assert Q[k] == v, k


def test_vertex_normals_ply_and_colormap(tmp_path):
    from coma_b200.cli.io import jet_rgb, read_point_cloud_ply, vertex_normals, write_point_cloud_ply
    # octahedron: vertex normals point radially
    v = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], dtype=np.float64)
    f = np.array([[0, 2, 4], [2, 1, 4], [1, 3, 4], [3, 0, 4], [2, 0, 5], [1, 2, 5], [3, 1, 5], [0, 3, 5]])
    np.testing.assert_allclose(vertex_normals(v, f), v, atol=1e-12)
    assert vertex_normals(np.vstack([v, [[9, 9, 9]]]), f)[-1].tolist() == [0, 0, 1]    # isolated vertex
    np.testing.assert_allclose(jet_rgb([0.0, 1.0]), [[0, 0, 0.5], [0.5, 0, 0]], atol=1e-12)
    np.testing.assert_allclose(jet_rgb([0.5])[0], [0.4838709677, 1.0, 0.4838709677], atol=1e-6)
    pth = str(tmp_path / "a" / "pc.ply")
    write_point_cloud_ply(pth, v, v, jet_rgb(np.linspace(0, 1, 6)))
    rec = read_point_cloud_ply(pth)
    assert rec.dtype.names == ("x", "y", "z", "nx", "ny", "nz", "red", "green", "blue") and len(rec) == 6
    np.testing.assert_array_equal(np.stack([rec["x"], rec["y"], rec["z"]], -1), v)
    assert rec["blue"][0] == 127 and rec["red"][-1] == 127


def _make_tree(root, S=5, H=40, O=12, seed=0):
    from coma_b200 import synth
    rng = np.random.default_rng(seed)
    V = 60
    samples = synth.make_samples(S, V, O, seed)
    faces = np.stack([rng.permutation(V)[:3] for _ in range(3 * V)])
    hidx = np.sort(rng.permutation(V)[:H])
    os.makedirs(f"{root}/mesh")
    pickle.dump(dict(N=H, N_raw=H, downsample_indices=hidx), open(f"{root}/mesh/smplx_star_downsampled_{H}.pickle", "wb"))
    ov, on = samples[0]["obj_verts"], samples[0]["obj_normals"]
    os.makedirs(f"{root}/asset_downsample/BEHAVE/backpack")
    pickle.dump(dict(obj_vertices_original=ov, obj_vertex_normals_original=on, obj_faces_original=np.zeros((1, 3), int),
                     downsample_indices=np.arange(O), N=O, N_raw=O, downsampled_pcd_points_raw=ov, downsampled_pcd_normal_raw=on),
                open(f"{root}/asset_downsample/BEHAVE/backpack/asset0_{O}.pickle", "wb"))
    for i, s in enumerate(samples):
        d = f"{root}/human_sample/BEHAVE/backpack/asset0/view:00001/mask:00002/carrying a backpack, full body"
        os.makedirs(d, exist_ok=True)
        pickle.dump(dict(verts=s["human_verts"], faces=faces, IoU=0.9, interscetion_ratio=0.0, num_inliers=20, z_min=0.0),
                    open(f"{d}/{i:06}.pickle", "wb"))
    d = f"{root}/human_sample/BEHAVE/backpack/asset0/view:00001/mask:00002/carrying a backpack, full body"
    return samples, faces, hidx, d


@pytest.mark.gpu
def test_extract_and_inference_end_to_end(tmp_path):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from constants.coma.qual import QUAL_AFFORDANCE_EXTRACTION_HYPERPARAMS_DICT as Q
    from coma_b200.cli.extract import inference, run_affordance_extraction
    from coma_b200.cli.io import read_point_cloud_ply, vertex_normals
    from oracle import oracle
    root = str(tmp_path)
    H, O = 40, 12
    samples, faces, hidx, sample_dir = _make_tree(root, S=5, H=H, O=O)
    pickle.dump("TOO LITTLE INLIERS", open(f"{sample_dir}/000099.pickle", "wb"))     # sentinel-string samples are skipped
    base = dict(human_res=str(H), object_res=str(O), enable_postfilter=False, spatial_grid_thres=0.24)
    common = dict(supercategories=["behave"], categories=["backpack"], prompts=None, camera_dir=f"{root}/cameras",
                  human_params_dir=f"{root}/human_preds", asset_downsample_dir=f"{root}/asset_downsample",
                  human_postfilter_dir=f"{root}/postfilter", human_sample_dir=f"{root}/human_sample",
                  coma_save_dir=f"{root}/extracted", affordance_save_dir=f"{root}/affordance", smplx_downsample_dir=f"{root}/mesh")
    outs = {}
    for key, fname in (("qual:backpack_human_contact", "human_contact.npy"), ("qual:backpack_object_contact", "object_contact.ply"),
                       ("qual:backpack_orientation", "orientational_tendency.npy"), ("qual:backpack_occupancy", "occupancy.npy")):
        hp = {**Q[key], **base}
        if key.endswith("occupancy"):
            hp["spatial_res"] = 10
        run_affordance_extraction(hyperparams=hp, hyperparams_key=key, skip_done=False, **common)
        main = "carrying a backpack"
        out = f"{root}/affordance/BEHAVE/backpack/asset0/{key}:{main}/{fname}"
        assert os.path.exists(out), out
        assert os.path.exists(f"{root}/extracted/BEHAVE/backpack/asset0/{key}:{main}.pickle")
        assert os.path.exists(f"{root}/extracted/BEHAVE/backpack/asset0/{key}:{main}.json")
        outs[key] = (out, hp)

    # accumulators in the exported pickle == oracle on the same ingest (vertex normals, index gather, fp32 rounding)
    exp = pickle.load(open(f"{root}/extracted/BEHAVE/backpack/asset0/qual:backpack_human_contact:carrying a backpack.pickle", "rb"))
    hv = np.stack([s["human_verts"][hidx] for s in samples])
    ov = np.stack([s["obj_verts"] for s in samples])
    cnt, nom = oracle.pair_accumulate(hv, ov, 0.24, 0.07, sum_order="cuda")   # ComA's default: the CUDA reference's association
    assert exp["used_count"] == 5
    np.testing.assert_array_equal(exp["significant_contact_count"], cnt)
    np.testing.assert_allclose(exp["contact_dist_expectation_grid_nom"], nom, rtol=1e-4)
    hn = np.stack([vertex_normals(s["human_verts"], faces)[hidx] for s in samples])
    hn = hn / (np.linalg.norm(hn, axis=-1, keepdims=True) + 1e-10)
    on = np.stack([s["obj_normals"] / (np.linalg.norm(s["obj_normals"], axis=-1, keepdims=True) + 1e-8) for s in samples])
    PH, _ = oracle.orient_accumulate(hn, on, oracle.fibonacci_sphere(250), 0.25, 1e-10, sum_order="cuda")
    np.testing.assert_allclose(exp["prob_grid_canon_human_wrt_obj"], PH, rtol=1e-4, atol=5 * 2.0 ** -31)   # cone-limited K3

    hc = np.load(outs["qual:backpack_human_contact"][0])
    assert hc.shape == (H,) and hc.dtype == np.float32 and np.nanmax(hc) == 1.0
    ply = read_point_cloud_ply(outs["qual:backpack_object_contact"][0])
    assert len(ply) == O and "red" in ply.dtype.names
    occ = np.load(outs["qual:backpack_occupancy"][0], allow_pickle=True).item()
    assert occ["prob_field"].shape == (10, 10, 10) and set(occ["spatial_grid_metadata"]) >= {"voxel_size", "start_point", "N_x"}

    # inference.py path: load the pickle back and reproduce the same file
    key, (out, hp) = "qual:backpack_human_contact", outs["qual:backpack_human_contact"]
    inference("BEHAVE", "backpack", f"{root}/extracted/BEHAVE/backpack/asset0/{key}:carrying a backpack.pickle",
              f"{root}/mesh/smplx_star_downsampled_{H}.pickle", f"{root}/asset_downsample/BEHAVE/backpack/asset0_{O}.pickle",
              "aggr-human-contact", key, hp, f"{root}/output")
    np.testing.assert_allclose(np.load(f"{root}/output/BEHAVE/backpack/human_contact.npy"), hc, rtol=1e-5, equal_nan=True)
    # --skip_done reloads the checkpoint instead of re-aggregating
    run_affordance_extraction(hyperparams=hp, hyperparams_key=key, skip_done=True, **common)
