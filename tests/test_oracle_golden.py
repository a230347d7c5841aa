"""Pins the CPU oracle (oracle/) to vectors produced by the unmodified reference (tests/golden/make_golden.py).

Bit-exact: nearest-vertex indices, contact counts, occupancy hit counts, significant-vertex index lists.
1e-4 relative (the north-star tolerance for fp32 quantities): proximity sums, orientation grids, read-outs.
"""
import os

import numpy as np
import pytest

from oracle import oracle

RTOL = 1e-4


def _load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name + ".npz")))


def test_fibonacci_grid_matches_reference(golden_dir):
    g = _load(golden_dir, "contact_small")
    # the reference exports the grid as fp32 (export -> to_np_torch_recursive)
    np.testing.assert_array_equal(oracle.fibonacci_sphere(int(g["N"])).astype(np.float32), g["canon_normal_grid"])


def test_nearest_vertex_bit_exact(golden_dir):
    g = _load(golden_dir, "nearest_small")
    idx = oracle.nearest_vertex(g["pts"], g["verts"])
    np.testing.assert_array_equal(idx, g["idx"])
    assert idx[0] == 3 and idx[1] == 0  # ties resolve to the lowest index


@pytest.mark.parametrize("name", ["contact_small", "contact_sigma02", "cuda_contact_small", "cuda_contact_sigma02", "cuda_contact_boundary"])
def test_contact_accumulators(golden_dir, name):
    """`cuda_*` fixtures were written by the reference running with device="cuda" on a B200 (tests/golden/make_golden_cuda.py):
    they pin the oracle's sum_order="cuda" association ((x2+z2)+y2, ATen's CUDA reduction); the others its CPU association."""
    if not os.path.exists(os.path.join(golden_dir, name + ".npz")):
        pytest.skip(f"{name}.npz not generated yet")
    g = _load(golden_dir, name)
    order = "cuda" if name.startswith("cuda_") else "cpu"
    size, thres, sigma, eps, ratio = g["params"]
    count, nom = oracle.pair_accumulate(g["hv"], g["ov"], thres, size, sum_order=order)
    np.testing.assert_array_equal(count, g["count"])                       # integer histogram: bit-exact
    np.testing.assert_allclose(nom, g["nom"], rtol=RTOL, atol=0)
    grid = oracle.fibonacci_sphere(int(g["N"]))
    PH, PO = oracle.orient_accumulate(g["hn"], g["on"], grid, sigma, eps, sum_order=order)
    # pure relative tolerance down to the smallest normal fp32 (below that the fp32 store rounds to denormals/zero)
    np.testing.assert_allclose(PH, g["PH"], rtol=RTOL, atol=1e-37)
    np.testing.assert_allclose(PO, g["PO"], rtol=RTOL, atol=1e-37)


@pytest.mark.parametrize("name", ["contact_small", "contact_sigma02"])
def test_contact_readouts(golden_dir, name):
    g = _load(golden_dir, name)
    size, thres, sigma, eps, ratio = g["params"]
    grid = oracle.fibonacci_sphere(int(g["N"]))
    used = int(g["used_count"])
    sig = oracle.significant_pairs(g["count"], ratio, used)
    np.testing.assert_array_equal(sig, g["sig_pairs"])
    # the golden script called get_aggregated_contact twice, compute_contact_map once, entropy once: 4 in-place
    # normalisations of the grids in total before the entropy was taken (utils/coma.py:328-330)
    PHn, POn = g["PH"], g["PO"]
    PHn, POn = oracle.normalize_normals(PHn, eps), oracle.normalize_normals(POn, eps)
    cm_h = oracle.contact_map(PHn, grid, g["nom"], g["denom"])
    agg_h, idx_o = oracle.aggregate_contact(cm_h, sig, "human")
    np.testing.assert_allclose(agg_h, g["agg_human"], rtol=RTOL, atol=1e-12)
    np.testing.assert_array_equal(idx_o, g["sig_obj_idx"])
    PHn, POn = oracle.normalize_normals(PHn, eps), oracle.normalize_normals(POn, eps)
    cm_o = oracle.contact_map(POn, grid, g["nom"], g["denom"])
    agg_o, idx_h = oracle.aggregate_contact(cm_o, sig, "obj")
    np.testing.assert_allclose(agg_o, g["agg_obj"], rtol=RTOL, atol=1e-12)
    np.testing.assert_array_equal(idx_h, g["sig_human_idx"])
    PHn, POn = oracle.normalize_normals(PHn, eps), oracle.normalize_normals(POn, eps)
    np.testing.assert_allclose(oracle.contact_map(PHn, grid, g["nom"], g["denom"]), g["contact_map_human"], rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(oracle.contact_map(POn, grid, g["nom"], g["denom"]), g["contact_map_obj"], rtol=RTOL, atol=1e-12)
    PHn, POn = oracle.normalize_normals(PHn, eps), oracle.normalize_normals(POn, eps)
    # entropy: sums of ~250 q*log(q) terms whose quantisation (round(P*1e6)) can flip on a 1-ulp difference of P
    np.testing.assert_allclose(oracle.entropy_score(PHn), g["entropy_human"], rtol=RTOL, atol=2e-5)
    np.testing.assert_allclose(oracle.entropy_score(POn), g["entropy_obj"], rtol=RTOL, atol=2e-5)


@pytest.mark.parametrize("name", ["occupancy_small", "occupancy_s30"])
def test_occupancy_bit_exact(golden_dir, name):
    g = _load(golden_dir, name)
    Sg = int(g["Sg"])
    centers, voxel = oracle.voxel_centers(Sg)
    # the reference exports its fp64 grid as fp32 (export -> to_np_torch_recursive)
    np.testing.assert_array_equal(centers[0].astype(np.float32), g["spatial_grid"][0, :, 0, 0])
    np.testing.assert_array_equal(centers[1].astype(np.float32), g["spatial_grid"][1, 0, :, 0])
    np.testing.assert_array_equal(centers[2].astype(np.float32), g["spatial_grid"][2, 0, 0, :])
    assert voxel == float(g["voxel_size"]) and voxel * float(g["tol"]) == float(g["rel_dist_thres"])
    grids = oracle.occupancy_accumulate(g["hv"], g["ov"], Sg, float(g["tol"]))
    np.testing.assert_array_equal(grids, g["grids"])                       # integer hit counts: bit-exact
    field, _ = oracle.occupancy_field(grids)
    np.testing.assert_allclose(field, g["field"], rtol=1e-6, equal_nan=True)


def test_canonicalize_matches_golden_through_scores(golden_dir):
    """Antipodal / degenerate normals are in contact_small (make_adversarial_samples): the reflect branch
    (utils/coma.py:143-145,169) must have been taken for object vertex 0 and produce finite unit vectors."""
    g = _load(golden_dir, "contact_small")
    eps = float(g["params"][3])
    c = oracle.canonicalize(g["hn"][-1], g["on"][-1], eps=eps)
    assert np.isfinite(c).all()
    np.testing.assert_allclose(np.linalg.norm(c, axis=-1), 1.0, atol=1e-6)
    # b = (0,0,-1): replacer = 2 (a.sub_p) sub_p - a, with sub_p = (0,1,0)
    a = g["hn"][-1].astype(np.float32)
    a = a / (np.linalg.norm(a, axis=-1, keepdims=True) + np.float32(eps))
    expect = np.stack([-a[:, 0], a[:, 1], -a[:, 2]], -1)
    np.testing.assert_allclose(c[:, 0, :], expect, atol=1e-6)
