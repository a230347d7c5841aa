"""Multi-GPU check on real NCCL (run under torchrun on N GPUs of one box): the row-sharded ComA / ComA_Occupancy with the sample
exchange must reproduce a single-GPU run of the same classes (rank 0 computes it on its own device) — counts bit-exact, fp32 sums
to 1e-5 (the exchange concatenates samples rank-major, so the fp32 summation order differs).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dist_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coma_b200 import dist as cdist  # noqa: E402
from coma_b200 import synth  # noqa: E402
from utils.coma import ComA, get_aggregated_contact  # noqa: E402
from utils.coma_occupancy import ComA_Occupancy  # noqa: E402

rank, world, local = cdist.init_process_group("nccl")
dev = f"cuda:{local}"
H, O, N, S, SG = 523, 180, 250, 37, 30
PROX = dict(spatial_grid_size=0.15, spatial_grid_thres=0.3)
samples = [dict(human_verts=s["human_verts"], human_normals=s["human_normals"], obj_verts=s["obj_verts"], obj_normals=s["obj_normals"])
           for s in synth.make_samples(S, H, O, seed=11)]
mine = cdist.sample_shard(S, rank, world)


def run(human_slice, idx, exchange):
    c = ComA(H, O, N, 0, proximity_settings=PROX, normal_gaussian_sigma=0.25, eps=1e-10, device=dev, human_slice=human_slice)
    for i in idx:
        c.register_sample_to_cache(**samples[i])
    c.aggregate_all_samples(exchange=exchange)
    out = dict(export=c.export(), agg_h=get_aggregated_contact(c, "human", 0.1), agg_o=get_aggregated_contact(c, "obj", 0.1),
               ent=c.compute_nonphysical_response_sphere(1e6, "human")["human"], used=c.used_count)
    occ = ComA_Occupancy(3.0, H, O, 0, SG, device=dev, human_slice=human_slice)
    for i in idx:
        occ.register_sample_to_cache(**samples[i])
    occ.aggregate_all_samples(exchange=exchange)
    out["occ_export"] = occ.export()
    out["field"] = occ.return_aggregated_spatial_grids().cpu().numpy()
    return out


sh = run(cdist.human_slice(H, rank, world), mine, True)
ok = True
if rank == 0:
    ref = run(None, list(range(S)), False)
    e, r = sh["export"], ref["export"]
    assert sh["used"] == ref["used"] == S
    np.testing.assert_array_equal(e["significant_contact_count"], r["significant_contact_count"])
    for k in ("contact_dist_expectation_grid_nom", "prob_grid_canon_human_wrt_obj", "prob_grid_canon_obj_wrt_human"):
        np.testing.assert_allclose(e[k], r[k], rtol=1e-5, atol=S * 2.0 ** -30)
    for k in ("agg_h", "agg_o"):
        np.testing.assert_array_equal(sh[k][1], ref[k][1])
        np.testing.assert_allclose(sh[k][0], ref[k][0], rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(sh["ent"], ref["ent"], rtol=1e-4, atol=2e-5)
    np.testing.assert_array_equal(sh["occ_export"]["spatial_occupancy_grids"], ref["occ_export"]["spatial_occupancy_grids"])
    np.testing.assert_allclose(sh["field"], ref["field"], rtol=1e-6, equal_nan=True)
    assert r["significant_contact_count"].sum() > 0 and len(ref["agg_h"][1]) > 0
    print(f"dist_check ok: world {world}, row-sharded + exchange == single GPU (H={H} O={O} N={N} S={S}, occupancy {SG}^3)", flush=True)
else:
    assert sh["export"] is None and sh["occ_export"] is None
dist.barrier()
dist.destroy_process_group()
