"""The N > 1 paths on CPU with world_size = 2 (gloo): sample-sharded ComA + all-reduce(SUM), H-sharded occupancy +
all-reduce(MAX, NaN-propagating). Per-rank accumulators are filled by the oracle (there is no GPU here); what is under
test is the sharding arithmetic and the collectives of the product classes."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from coma_b200 import dist as cdist
    from coma_b200 import synth
    from oracle import oracle
    from utils.coma import ComA
    from utils.coma_occupancy import ComA_Occupancy
    cdist.init_process_group("gloo")
    H, O, N, S = 12, 7, 32, 9
    samples = synth.make_samples(S, H, O, seed=5)
    mine = cdist.sample_shard(S, rank, world)
    c = ComA(H, O, N, 0, proximity_settings=dict(spatial_grid_size=0.07, spatial_grid_thres=0.24), normal_gaussian_sigma=0.25,
             eps=1e-10, device="cpu")
    hv = np.stack([samples[i]["human_verts"] for i in mine]); hn = np.stack([samples[i]["human_normals"] for i in mine])
    ov = np.stack([samples[i]["obj_verts"] for i in mine]); on = np.stack([samples[i]["obj_normals"] for i in mine])
    cnt, nom = oracle.pair_accumulate(hv, ov, 0.24, 0.07)
    PH, PO = oracle.orient_accumulate(hn, on, oracle.fibonacci_sphere(N), 0.25, 1e-10)
    c.significant_contact_count += torch.from_numpy(cnt)
    c.contact_dist_expectation_grid_nom += torch.from_numpy(nom)
    c.contact_dist_expectation_grid_denom += float(len(mine))
    c.prob_grid_canon_human_wrt_obj += torch.from_numpy(PH)
    c.prob_grid_canon_obj_wrt_human += torch.from_numpy(PO)
    c.used_count = len(mine)
    c.all_reduce()
    e = c.export()

    # occupancy: H-sharded, every rank sees all samples
    Sg = 8
    h0, h1 = cdist.human_slice(H, rank, world)
    occ = ComA_Occupancy(3.0, H, O, 0, Sg, device="cpu", human_slice=(h0, h1))
    assert occ.spatial_occupancy_grids.shape[0] == h1 - h0
    hv_all = np.stack([s["human_verts"] for s in samples]); ov_all = np.stack([s["obj_verts"] for s in samples])
    hv_all[:, 0] = 50.0     # vertex 0 never hits -> NaN row on rank 0 must reach every rank through the MAX all-reduce
    full = oracle.occupancy_accumulate(hv_all, ov_all, Sg, 3.0)
    occ.spatial_occupancy_grids += torch.from_numpy(full[h0:h1])
    _, norm = oracle.occupancy_field(full[h0:h1])
    local = np.where(np.isnan(norm).any(0), np.nan, norm.max(0)).astype(np.float32)
    from coma_b200.coma_occupancy import _all_reduce_max_nan
    field = _all_reduce_max_nan(torch.from_numpy(local)).numpy()
    if rank == 0:
        np.savez(os.path.join(out_dir, "r0.npz"), count=e["significant_contact_count"], nom=e["contact_dist_expectation_grid_nom"],
                 denom=e["contact_dist_expectation_grid_denom"], PH=e["prob_grid_canon_human_wrt_obj"], used=e["used_count"], field=field)
    dist.barrier()
    dist.destroy_process_group()


def test_world2_gloo(tmp_path):
    from coma_b200 import synth
    from oracle import oracle
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r = np.load(tmp_path / "r0.npz")
    H, O, N, S = 12, 7, 32, 9
    samples = synth.make_samples(S, H, O, seed=5)
    hv = np.stack([s["human_verts"] for s in samples]); hn = np.stack([s["human_normals"] for s in samples])
    ov = np.stack([s["obj_verts"] for s in samples]); on = np.stack([s["obj_normals"] for s in samples])
    cnt, nom = oracle.pair_accumulate(hv, ov, 0.24, 0.07)
    PH, _ = oracle.orient_accumulate(hn, on, oracle.fibonacci_sphere(N), 0.25, 1e-10)
    np.testing.assert_array_equal(r["count"], cnt)            # integer counts survive the SUM exactly
    assert int(r["used"]) == S and (r["denom"] == S).all()
    np.testing.assert_allclose(r["nom"], nom, rtol=1e-5)
    np.testing.assert_allclose(r["PH"], PH, rtol=1e-5, atol=1e-30)
    hv[:, 0] = 50.0
    full = oracle.occupancy_accumulate(hv, ov, 8, 3.0)
    ref_field, _ = oracle.occupancy_field(full)
    assert np.isnan(ref_field).all()                           # a never-hit vertex poisons the reference's max (NaN)
    np.testing.assert_array_equal(np.isnan(r["field"]), np.isnan(ref_field))
