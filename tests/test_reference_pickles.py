"""Checkpoints written BY THE REFERENCE (tests/golden/ref_*.pickle, produced by tests/golden/make_golden.py running the
unmodified snuvclab/coma classes) load through the drop-in classes — the `--skip_done` / `src/coma/inference.py` path — and
the reverse direction: a pickle exported by the drop-in class loads in the reference (GPU box: oracle/_ref)."""
import os
import pickle

import numpy as np
import pytest
import torch


def test_reference_pickle_layout_matches_export_keys(golden_dir):
    """CPU: unpickling needs an importable `utils.coma.negative_exp` (pickled by reference) — the shim provides it — and the
    key set / dtypes equal what the drop-in classes export."""
    from coma_b200.coma import _EXPORT_KEYS
    from coma_b200.coma_occupancy import _EXPORT_KEYS as OCC_KEYS
    import utils.coma as shim
    d = pickle.load(open(os.path.join(golden_dir, "ref_coma_sigma02.pickle"), "rb"))
    assert set(d) == set(_EXPORT_KEYS)
    assert d["contact_dist_func"].func is shim.negative_exp and d["contact_dist_func"].keywords == d["proximity_settings"]
    assert d["prob_grid_canon_human_wrt_obj"].dtype == np.float32 and d["canon_normal_grid"].dtype == np.float32
    o = pickle.load(open(os.path.join(golden_dir, "ref_occupancy_small.pickle"), "rb"))
    assert set(o) == set(OCC_KEYS)
    assert o["spatial_indexgrid"].dtype == np.int64 and o["spatial_occupancy_grids"].dtype == np.float32


@pytest.mark.gpu
def test_load_reference_written_pickles_and_read_out(golden_dir, tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from utils.coma import ComA, get_aggregated_contact
    from utils.coma_occupancy import ComA_Occupancy
    g = dict(np.load(os.path.join(golden_dir, "contact_sigma02.npz")))
    size, thres, sigma, eps, ratio = (float(v) for v in g["params"])
    S, H, _ = g["hv"].shape
    O, N = g["ov"].shape[1], int(g["N"])
    c = ComA(H, O, N, 0, proximity_settings=dict(spatial_grid_size=size, spatial_grid_thres=thres), normal_gaussian_sigma=sigma, eps=eps, device="cuda")
    c.load(os.path.join(golden_dir, "ref_coma_sigma02.pickle"))
    c.device = "cuda"
    assert c.used_count == S and c.prob_grid_canon_human_wrt_obj.is_cuda
    np.testing.assert_array_equal(c.significant_contact_count.cpu().numpy(), g["count"])
    agg_h, idx_o = get_aggregated_contact(c, "human", ratio)
    agg_o, idx_h = get_aggregated_contact(c, "obj", ratio)
    np.testing.assert_array_equal(idx_o, g["sig_obj_idx"])
    np.testing.assert_array_equal(idx_h, g["sig_human_idx"])
    np.testing.assert_allclose(agg_h, g["agg_human"], rtol=1e-4, atol=1e-12)
    np.testing.assert_allclose(agg_o, g["agg_obj"], rtol=1e-4, atol=1e-12)
    ent = c.compute_nonphysical_response_sphere(n_bin=1e6, nonphysical_type="human")["human"]
    np.testing.assert_allclose(ent, g["entropy_human"], rtol=1e-4, atol=2e-5)

    go = dict(np.load(os.path.join(golden_dir, "occupancy_small.npz")))
    occ = ComA_Occupancy(float(go["tol"]), go["hv"].shape[1], go["ov"].shape[1], 0, int(go["Sg"]), device="cuda")
    occ.load(os.path.join(golden_dir, "ref_occupancy_small.pickle"))
    np.testing.assert_array_equal(occ.spatial_occupancy_grids.cpu().numpy(), go["grids"])
    np.testing.assert_allclose(occ.return_aggregated_spatial_grids().cpu().numpy(), go["field"], rtol=1e-6, equal_nan=True)

    # reverse direction: our export -> the reference's load + read-out (reference staged in oracle/_ref)
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("oracle/_ref not staged")
    ref = ref_loader.load()
    c2 = ComA(H, O, N, 0, proximity_settings=dict(spatial_grid_size=size, spatial_grid_thres=thres), normal_gaussian_sigma=sigma, eps=eps, device="cuda")
    for s in range(S):
        c2.register_sample_to_cache(human_verts=g["hv"][s], human_normals=g["hn"][s], obj_verts=g["ov"][s], obj_normals=g["on"][s])
    c2.aggregate_all_samples()
    pth = str(tmp_path / "ours.pickle")
    c2.export(save_pth=pth)
    rc = ref.ComA(H, O, N, 0, proximity_settings=dict(spatial_grid_size=size, spatial_grid_thres=thres), normal_gaussian_sigma=sigma, eps=eps, device="cuda")
    rc.load(pth)
    r_agg, r_idx = ref.get_aggregated_contact(rc, "human", ratio)
    np.testing.assert_array_equal(r_idx, g["sig_obj_idx"])
    np.testing.assert_allclose(r_agg, g["agg_human"], rtol=1e-4, atol=1e-12)
