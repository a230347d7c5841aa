"""Row-sharded multi-GPU paths on CPU with world_size = 2 (gloo): `ComA(human_slice=...)` / `ComA_Occupancy(human_slice=...)`
with the sample exchange (`aggregate_all_samples(exchange=True)`), the collective read-outs, export / load of sharded
instances, and the CLI driver under two ranks (`run_affordance_extraction` + `write_affordance`, the path the round-1 advisor
found dead-locking). The kernels are replaced by the oracle (tests/_cpu_shim.py — there is no GPU here); what is under test
is the host-side sharding / collective logic, against a single-process run of the same classes."""
import os
import pickle
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

H, O, N, S, SG = 13, 7, 32, 9, 8
PROX = dict(spatial_grid_size=0.07, spatial_grid_thres=0.45)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run_classes(samples, mine, human_slice, exchange, out):
    from utils.coma import ComA, get_aggregated_contact
    from utils.coma_occupancy import ComA_Occupancy
    c = ComA(H, O, N, 0, proximity_settings=PROX, normal_gaussian_sigma=0.25, eps=1e-10, device="cpu", human_slice=human_slice)
    for i in mine:
        c.register_sample_to_cache(**samples[i])
    c.aggregate_all_samples(exchange=exchange)
    out["used"] = c.used_count
    out["export"] = c.export()
    out["agg_h"], out["idx_o"] = get_aggregated_contact(c, "human", 0.12)
    out["agg_o"], out["idx_h"] = get_aggregated_contact(c, "obj", 0.12)
    out["sig"] = c.significant_contact_pairs(0.12)
    out["cm"] = c.compute_contact_map("both")
    out["ent"] = c.compute_nonphysical_response_sphere(1e6, "both")
    occ = ComA_Occupancy(3.0, H, O, 0, SG, device="cpu", human_slice=human_slice)
    for i in mine:
        occ.register_sample_to_cache(**samples[i])
    occ.aggregate_all_samples(exchange=exchange)
    out["occ_used"] = occ.used_count
    out["occ_export"] = occ.export()
    out["field_sel"] = occ.return_aggregated_spatial_grids(human_indices=[1, 2, 4]).numpy()   # all on rank 0 of 2: rank 1 selects nothing
    out["field"] = occ.return_aggregated_spatial_grids().numpy()
    return c, occ


def _samples():
    from coma_b200 import synth
    return [dict(human_verts=s["human_verts"], human_normals=s["human_normals"], obj_verts=s["obj_verts"], obj_normals=s["obj_normals"])
            for s in synth.make_samples(S, H, O, seed=5)]


def _tree_args(root):
    return dict(supercategories=["behave"], categories=["backpack"], prompts=None, camera_dir=f"{root}/cameras",
                human_params_dir=f"{root}/human_preds", asset_downsample_dir=f"{root}/asset_downsample",
                human_postfilter_dir=f"{root}/postfilter", human_sample_dir=f"{root}/human_sample",
                smplx_downsample_dir=f"{root}/mesh", device="cpu")


KEYS = (("qual:backpack_human_contact", "human_contact.npy"), ("qual:backpack_object_contact", "object_contact.ply"),
        ("qual:backpack_orientation", "orientational_tendency.npy"), ("qual:backpack_occupancy", "occupancy.npy"))


def _run_cli(root, tag):
    from constants.coma.qual import QUAL_AFFORDANCE_EXTRACTION_HYPERPARAMS_DICT as Q
    from coma_b200.cli.extract import run_affordance_extraction
    for key, _ in KEYS:
        hp = {**Q[key], **dict(human_res="40", object_res="12", enable_postfilter=False, spatial_grid_thres=0.24)}
        if key.endswith("occupancy"):
            hp["spatial_res"] = 10
        run_affordance_extraction(hyperparams=hp, hyperparams_key=key, skip_done=False, coma_save_dir=f"{root}/extracted_{tag}",
                                  affordance_save_dir=f"{root}/affordance_{tag}", **_tree_args(root))


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from coma_b200 import dist as cdist
    from tests import _cpu_shim
    _cpu_shim.install()
    cdist.init_process_group("gloo")
    samples = _samples()
    out = {}
    c, occ = _run_classes(samples, cdist.sample_shard(S, rank, world), cdist.human_slice(H, rank, world), True, out)
    assert c.prob_grid_canon_human_wrt_obj.shape[0] == cdist.human_slice(H, rank, world)[1] - cdist.human_slice(H, rank, world)[0]
    if rank == 0:
        pickle.dump(out["export"], open(f"{out_dir}/coma.pickle", "wb"))
    dist.barrier()
    # a sharded instance loads the full pickle and keeps its rows
    from utils.coma import ComA, get_aggregated_contact
    again = ComA(H, O, N, 0, proximity_settings=PROX, normal_gaussian_sigma=0.25, eps=1e-10, device="cpu",
                 human_slice=cdist.human_slice(H, rank, world))
    again.load(f"{out_dir}/coma.pickle")
    out["agg_h_reloaded"], _ = get_aggregated_contact(again, "human", 0.12)
    assert (out["export"] is None) == (rank != 0) and (out["occ_export"] is None) == (rank != 0)
    pickle.dump(out, open(f"{out_dir}/rank{rank}.pickle", "wb"))
    _run_cli(out_dir, "w2")       # every rank enters the collective read-outs; only rank 0 writes
    dist.barrier()
    dist.destroy_process_group()


import pytest


@pytest.fixture
def cpu_kernels():
    from tests import _cpu_shim
    restore = _cpu_shim.install()
    yield
    restore()


def test_row_sharded_world2_equals_single_process(tmp_path, cpu_kernels):
    from tests.test_cli import _make_tree
    root = str(tmp_path)
    _make_tree(root, S=5, H=40, O=12)
    port = _free_port()
    mp.spawn(_worker, args=(2, port, root), nprocs=2, join=True)
    ref = {}
    _run_classes(_samples(), list(range(S)), None, False, ref)
    r0, r1 = (pickle.load(open(f"{root}/rank{r}.pickle", "rb")) for r in (0, 1))
    assert r0["used"] == r1["used"] == ref["used"] == S and r0["occ_used"] == S
    e, re_ = r0["export"], ref["export"]
    assert set(e) == set(re_) and e["human_res"] == H
    np.testing.assert_array_equal(e["significant_contact_count"], re_["significant_contact_count"])
    np.testing.assert_array_equal(e["contact_dist_expectation_grid_denom"], re_["contact_dist_expectation_grid_denom"])
    for k in ("contact_dist_expectation_grid_nom", "prob_grid_canon_human_wrt_obj", "prob_grid_canon_obj_wrt_human"):
        assert e[k].shape == re_[k].shape
        np.testing.assert_allclose(e[k], re_[k], rtol=1e-5, atol=1e-30)     # sample order differs (rank-major) -> fp32 sum order
    for r in (r0, r1):                                                       # every rank returns the FULL read-outs
        np.testing.assert_array_equal(r["idx_o"], ref["idx_o"])
        np.testing.assert_array_equal(r["idx_h"], ref["idx_h"])
        np.testing.assert_array_equal(r["sig"], ref["sig"])
        for k in ("agg_h", "agg_o"):
            assert r[k].shape == ref[k].shape
            np.testing.assert_allclose(r[k], ref[k], rtol=1e-5, atol=1e-12)
        np.testing.assert_allclose(r["agg_h_reloaded"], ref["agg_h"], rtol=1e-5, atol=1e-12)
        for k in ("human", "obj"):
            np.testing.assert_allclose(r["cm"][k], ref["cm"][k], rtol=1e-5, atol=1e-12)
            np.testing.assert_allclose(r["ent"][k], ref["ent"][k], rtol=1e-4, atol=2e-5)
        np.testing.assert_allclose(r["field"], ref["field"], rtol=1e-6, equal_nan=True)
        np.testing.assert_allclose(r["field_sel"], ref["field_sel"], rtol=1e-6, equal_nan=True)
    assert ref["idx_o"].size > 0 and ref["idx_h"].size > 0
    np.testing.assert_array_equal(r0["occ_export"]["spatial_occupancy_grids"], ref["occ_export"]["spatial_occupancy_grids"])

    # CLI under two ranks == CLI in one process (same files, same contents)
    _run_cli(root, "w1")
    for key, fname in KEYS:
        a = f"{root}/affordance_w2/BEHAVE/backpack/asset0/{key}:carrying a backpack/{fname}"
        b = f"{root}/affordance_w1/BEHAVE/backpack/asset0/{key}:carrying a backpack/{fname}"
        assert os.path.exists(a) and os.path.exists(b), (a, b)
        if fname.endswith(".npy"):
            x, y = np.load(a, allow_pickle=True), np.load(b, allow_pickle=True)
            if x.dtype == object:
                np.testing.assert_allclose(x.item()["prob_field"], y.item()["prob_field"], rtol=1e-6, equal_nan=True)
            else:
                np.testing.assert_allclose(x, y, rtol=1e-4, atol=2e-5, equal_nan=True)
        pa = pickle.load(open(f"{root}/extracted_w2/BEHAVE/backpack/asset0/{key}:carrying a backpack.pickle", "rb"))
        pb = pickle.load(open(f"{root}/extracted_w1/BEHAVE/backpack/asset0/{key}:carrying a backpack.pickle", "rb"))
        assert pa["used_count"] == pb["used_count"] == 5
        k = "spatial_occupancy_grids" if key.endswith("occupancy") else "significant_contact_count"
        np.testing.assert_array_equal(pa[k], pb[k])


def test_empty_selection_raises_like_reference(cpu_kernels):
    from utils.coma_occupancy import ComA_Occupancy
    occ = ComA_Occupancy(3.0, 4, 3, 0, 6, device="cpu")
    with pytest.raises(IndexError):
        occ.return_aggregated_spatial_grids(human_indices=[])
