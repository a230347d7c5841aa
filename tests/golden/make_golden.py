"""Generate the committed golden fixtures by running the UNMODIFIED reference (snuvclab/coma @ /root/reference)
on seeded synthetic inputs.  Run in the dev container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

The reference is pure Python; it imports after stubbing open3d / trimesh / easydict.
Outputs: tests/golden/*.npz.  The reference has no tests / golden vectors of its own, so these
reference-generated vectors are the parity pins for the ComA path.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("COMA_REFERENCE", "/root/reference")

for m in ("open3d", "trimesh", "easydict"):
    sys.modules[m] = types.ModuleType(m)
sys.modules["easydict"].EasyDict = dict
# the repo ships a pickle-compat `utils` shim; the reference's `utils` must win here
sys.path.insert(0, REF)
import torch  # noqa: E402
from utils.coma import ComA, get_aggregated_contact  # noqa: E402
from utils.coma_occupancy import ComA_Occupancy  # noqa: E402

import importlib.util  # noqa: E402
_spec = importlib.util.spec_from_file_location("coma_b200_synth", os.path.join(ROOT, "coma_b200", "synth.py"))
synth = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(synth)

torch.set_num_threads(8)


def contact_case(name, H, O, N, S, size, thres, sigma, eps, ratio, adversarial, seed):
    samples = synth.make_samples(S, H, O, seed)
    if adversarial:
        samples = samples + synth.make_adversarial_samples(H, O, thres, seed + 1)
        # keep the object fixed across ALL samples
        for s in samples:
            s["obj_verts"] = samples[-1]["obj_verts"].copy()
            s["obj_normals"] = samples[-1]["obj_normals"].copy()
    coma = ComA(human_res=H, obj_res=O, normal_res=N, spatial_res=0,
                proximity_settings=dict(spatial_grid_size=size, spatial_grid_thres=thres),
                normal_gaussian_sigma=sigma, eps=eps, device="cpu")
    for s in samples:
        coma.register_sample_to_cache(**{k: v.copy() for k, v in s.items()})
    coma.aggregate_all_samples()
    exp = coma.export()
    out = dict(
        hv=np.stack([s["human_verts"] for s in samples]), hn=np.stack([s["human_normals"] for s in samples]),
        ov=np.stack([s["obj_verts"] for s in samples]), on=np.stack([s["obj_normals"] for s in samples]),
        params=np.array([size, thres, sigma, eps, ratio], dtype=np.float64), N=np.int64(N),
        canon_normal_grid=exp["canon_normal_grid"],
        PH=exp["prob_grid_canon_human_wrt_obj"], PO=exp["prob_grid_canon_obj_wrt_human"],
        nom=exp["contact_dist_expectation_grid_nom"], denom=exp["contact_dist_expectation_grid_denom"],
        count=exp["significant_contact_count"], used_count=np.int64(exp["used_count"]),
    )
    # read-outs (each call re-normalises the grids in place, exactly like the scripts do)
    agg_h, idx_o = get_aggregated_contact(coma, "human", ratio)
    agg_o, idx_h = get_aggregated_contact(coma, "obj", ratio)
    cm = coma.compute_contact_map("both", as_numpy=True)
    ent = coma.compute_nonphysical_response_sphere(n_bin=1e6, nonphysical_type="both", as_numpy=True)
    out.update(agg_human=agg_h, sig_obj_idx=idx_o, agg_obj=agg_o, sig_human_idx=idx_h,
               contact_map_human=cm["human"], contact_map_obj=cm["obj"],
               entropy_human=ent["human"], entropy_obj=ent["obj"],
               sig_pairs=coma.significant_contact_pairs(ratio))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "count sum", out["count"].sum(), "PH max", out["PH"].max(), "sig", out["sig_pairs"].sum())


def pickle_case():
    """Checkpoints WRITTEN BY THE REFERENCE (`export(save_pth)`, utils/coma.py:582-597 / utils/coma_occupancy.py:314-330) for
    the contact_sigma02 / occupancy_small cases: the drop-in classes must load them (`--skip_done`, src/coma/inference.py) —
    including the `functools.partial(utils.coma.negative_exp)` the reference pickles by reference."""
    g = dict(np.load(os.path.join(HERE, "contact_sigma02.npz")))
    size, thres, sigma, eps, ratio = (float(v) for v in g["params"])
    S, H, _ = g["hv"].shape
    coma = ComA(human_res=H, obj_res=g["ov"].shape[1], normal_res=int(g["N"]), spatial_res=0,
                proximity_settings=dict(spatial_grid_size=size, spatial_grid_thres=thres), normal_gaussian_sigma=sigma, eps=eps, device="cpu")
    for s in range(S):
        coma.register_sample_to_cache(human_verts=g["hv"][s].copy(), human_normals=g["hn"][s].copy(), obj_verts=g["ov"][s].copy(),
                                      obj_normals=g["on"][s].copy())
    coma.aggregate_all_samples()
    coma.export(save_pth=os.path.join(HERE, "ref_coma_sigma02.pickle"))
    g = dict(np.load(os.path.join(HERE, "occupancy_small.npz")))
    S, H, _ = g["hv"].shape
    occ = ComA_Occupancy(scale_tolerance=float(g["tol"]), human_res=H, obj_res=g["ov"].shape[1], normal_res=0, spatial_res=int(g["Sg"]), device="cpu")
    for s in range(S):
        occ.register_sample_to_cache(human_verts=g["hv"][s].copy(), human_normals=g["hn"][s].copy(), obj_verts=g["ov"][s].copy(),
                                     obj_normals=g["on"][s].copy())
    occ.aggregate_all_samples()
    occ.export(save_pth=os.path.join(HERE, "ref_occupancy_small.pickle"))
    print("reference-written pickles:", os.path.getsize(os.path.join(HERE, "ref_coma_sigma02.pickle")),
          os.path.getsize(os.path.join(HERE, "ref_occupancy_small.pickle")), "bytes")


def occupancy_case(name, H, O, Sg, S, tol, seed):
    samples = synth.make_samples(S, H, O, seed)
    occ = ComA_Occupancy(scale_tolerance=tol, human_res=H, obj_res=O, normal_res=0, spatial_res=Sg, device="cpu")
    for s in samples:
        occ.register_sample_to_cache(**{k: v.copy() for k, v in s.items()})
    occ.aggregate_all_samples()
    exp = occ.export()
    grids = exp["spatial_occupancy_grids"].copy()
    field = occ.return_aggregated_spatial_grids(human_indices=None).cpu().numpy()
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        hv=np.stack([s["human_verts"] for s in samples]), hn=np.stack([s["human_normals"] for s in samples]),
        ov=np.stack([s["obj_verts"] for s in samples]), on=np.stack([s["obj_normals"] for s in samples]),
        Sg=np.int64(Sg), tol=np.float64(tol), grids=grids, field=field,
        spatial_grid=exp["spatial_grid"], start_point=exp["spatial_grid_metadata"]["start_point"],
        voxel_size=np.float64(exp["spatial_grid_metadata"]["voxel_size"]), rel_dist_thres=np.float64(exp["rel_dist_thres"]),
    )
    print(name, "hits", grids.sum(), "per vertex-sample", grids.sum() / (H * S))


def nearest_case(name, V, N, seed):
    rng = np.random.default_rng(seed)
    verts = rng.standard_normal((V, 3))
    pts = verts[rng.integers(0, V, N)] + rng.standard_normal((N, 3)) * 0.05
    verts[V // 2] = verts[3]          # exact duplicates -> argmin ties resolve to the lowest index
    verts[V - 1] = verts[0]
    pts[0] = verts[3]
    pts[1] = verts[0]
    pts[2] = 0.5 * (verts[5] + verts[6])  # equidistant in exact arithmetic
    # the two lines of utils/coma.py:90-91, verbatim semantics
    squared_dists = np.sum(np.square(pts[None, :, :] - verts[:, None, :]), axis=-1)
    idx = np.argmin(squared_dists, axis=0)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), pts=pts, verts=verts, idx=idx.astype(np.int64))
    print(name, idx[:4])


if __name__ == "__main__":
    if "--pickles-only" in sys.argv:
        pickle_case()
        sys.exit(0)
    contact_case("contact_small", H=24, O=12, N=250, S=5, size=0.07, thres=0.03, sigma=0.25, eps=1e-10, ratio=0.1,
                 adversarial=True, seed=42)
    contact_case("contact_sigma02", H=20, O=9, N=64, S=4, size=0.06, thres=0.24, sigma=0.2, eps=1e-10, ratio=0.3,
                 adversarial=False, seed=3)
    occupancy_case("occupancy_small", H=40, O=6, Sg=12, S=4, tol=3.0, seed=5)
    occupancy_case("occupancy_s30", H=6, O=4, Sg=30, S=3, tol=3.0, seed=6)
    nearest_case("nearest_small", V=700, N=96, seed=11)
    pickle_case()
