"""The UNMODIFIED reference run with device="cuda" on the B200 (its production setting, src/coma/extract_coma.py:329,346)
as the parity anchor: (1) does the reference itself agree CPU-vs-CUDA on the bit-exact quantities for the adversarial
threshold set (SURVEY §7 "which oracle is the reference"); (2) the kernels against the CUDA reference.
The reference modules are staged by oracle/make_ref.py into git-ignored oracle/_ref/ (they travel with gpurun)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("oracle/_ref not staged (python oracle/make_ref.py in the dev container)")
    return ref_loader.load()


def _case(H, O, S, thres, seed):
    from coma_b200 import synth
    samples = synth.make_samples(S, H, O, seed=seed) + synth.make_adversarial_samples(H, O, thres, seed=seed + 1)
    for s in samples:   # the occupancy class asserts a fixed object
        s["obj_verts"], s["obj_normals"] = samples[-1]["obj_verts"].copy(), samples[-1]["obj_normals"].copy()
    return samples


def _run(cls, samples, **kw):
    c = cls(**kw)
    for s in samples:
        c.register_sample_to_cache(**{k: v.copy() for k, v in s.items()})
    c.aggregate_all_samples()
    return c


@pytest.mark.parametrize("H,O,N,S,thres,sigma", [(96, 40, 250, 6, 0.05, 0.25), (64, 36, 250, 5, 0.03, 0.2), (33, 20, 64, 3, 0.24, 0.1)])
def test_contact_reference_cpu_vs_cuda_vs_kernels(ref, H, O, N, S, thres, sigma):
    from utils.coma import ComA, get_aggregated_contact
    samples = _case(H, O, S, thres, seed=H)
    kw = dict(human_res=H, obj_res=O, normal_res=N, spatial_res=0, proximity_settings=dict(spatial_grid_size=0.15, spatial_grid_thres=thres),
              normal_gaussian_sigma=sigma, eps=1e-10)
    r_cpu = _run(ref.ComA, samples, device="cpu", **kw).export()
    rc = _run(ref.ComA, samples, device="cuda", **kw)
    r_cuda = rc.export()
    mine_c = _run(ComA, samples, device="cuda", **kw)
    mine = mine_c.export()
    # (1) the reference against itself. ATen adds the 3-term sums as (x2+z2)+y2 on CUDA and (x2+y2)+z2 on the CPU, so the two runs
    # of the SAME reference may disagree on `count` for pairs within an ulp of the threshold (tests/golden/cuda_contact_boundary.npz
    # holds such pairs) and agree on the fp32 grids only to ~3e-5 relative (canonical normals differ in the last bit, amplified by
    # 1/sigma^2). The drop-in class follows the CUDA run (ComA.reference_sum_order = "cuda"); "cpu" reproduces the CPU run.
    np.testing.assert_allclose(r_cpu["prob_grid_canon_human_wrt_obj"], r_cuda["prob_grid_canon_human_wrt_obj"], rtol=1e-4, atol=1e-30)
    mine_cpu = _run(type("ComACpuOrder", (ComA,), dict(reference_sum_order="cpu")), samples, device="cuda", **kw).export()
    np.testing.assert_array_equal(mine_cpu["significant_contact_count"], r_cpu["significant_contact_count"])
    # (2) kernels against the CUDA reference
    np.testing.assert_array_equal(mine["significant_contact_count"], r_cuda["significant_contact_count"])
    np.testing.assert_array_equal(mine["contact_dist_expectation_grid_denom"], r_cuda["contact_dist_expectation_grid_denom"])
    np.testing.assert_allclose(mine["contact_dist_expectation_grid_nom"], r_cuda["contact_dist_expectation_grid_nom"], rtol=1e-4)
    for k in ("prob_grid_canon_human_wrt_obj", "prob_grid_canon_obj_wrt_human"):
        a, b = mine[k], r_cuda[k]
        tol = 1e-4 * np.abs(b) + 1e-30 + len(samples) * 2.0 ** -31      # cone-limited K3: terms < 2^-32 are dropped
        assert (np.abs(a.astype(np.float64) - b) <= tol).all(), k
    assert r_cuda["significant_contact_count"].sum() > 0
    for typ in ("human", "obj"):
        a, ia = get_aggregated_contact(mine_c, typ, 0.1)
        b, ib = ref.get_aggregated_contact(rc, typ, 0.1)
        np.testing.assert_array_equal(ia, ib)
        np.testing.assert_allclose(a, b, rtol=1e-4, atol=1e-12)


@pytest.mark.parametrize("H,O,Sg,S", [(64, 5, 30, 6), (40, 4, 12, 4)])
def test_occupancy_reference_cpu_vs_cuda_vs_kernels(ref, H, O, Sg, S):
    from utils.coma_occupancy import ComA_Occupancy
    samples = _case(H, O, S, 0.05, seed=Sg)
    kw = dict(scale_tolerance=3.0, human_res=H, obj_res=O, normal_res=0, spatial_res=Sg)
    r_cpu = _run(ref.ComA_Occupancy, samples, device="cpu", **kw).export()
    rc = _run(ref.ComA_Occupancy, samples, device="cuda", **kw)
    r_cuda = rc.export()
    mc = _run(ComA_Occupancy, samples, device="cuda", **kw)
    mine = mc.export()
    np.testing.assert_array_equal(r_cpu["spatial_occupancy_grids"], r_cuda["spatial_occupancy_grids"])
    np.testing.assert_array_equal(mine["spatial_occupancy_grids"], r_cuda["spatial_occupancy_grids"])
    assert r_cuda["spatial_occupancy_grids"].sum() > 0
    np.testing.assert_allclose(mc.return_aggregated_spatial_grids().cpu().numpy(), rc.return_aggregated_spatial_grids().cpu().numpy(),
                               rtol=1e-6, equal_nan=True)


def test_baseline_cfg1_in_full_vs_reference_cuda(ref):
    """BASELINE.json configs[0] at FULL size — 32 samples, 1000 human x 180 object vertices x 250 bins, preset
    qual:backpack_human_contact (constants/coma/qual.py: 0.07 / 0.03 / 0.25 / 1e-10) — the unmodified reference with device="cuda"
    against the drop-in class: counts bit-exact, fp32 accumulators and every read-out the scripts use at 1e-4."""
    from coma_b200 import synth
    from utils.coma import ComA, get_aggregated_contact, get_nonphysical_score
    H, O, N, S = 1000, 180, 250, 32
    samples = synth.make_samples(S, H, O, seed=1)
    kw = dict(human_res=H, obj_res=O, normal_res=N, spatial_res=0, proximity_settings=dict(spatial_grid_size=0.07, spatial_grid_thres=0.03),
              normal_gaussian_sigma=0.25, eps=1e-10)
    rc = _run(ref.ComA, samples, device="cuda", **kw)
    mc = _run(ComA, samples, device="cuda", **kw)
    r, m = rc.export(), mc.export()
    assert set(r.keys()) == set(m.keys())
    np.testing.assert_array_equal(m["significant_contact_count"], r["significant_contact_count"])
    np.testing.assert_array_equal(m["contact_dist_expectation_grid_denom"], r["contact_dist_expectation_grid_denom"])
    np.testing.assert_allclose(m["contact_dist_expectation_grid_nom"], r["contact_dist_expectation_grid_nom"], rtol=1e-4)
    for k in ("prob_grid_canon_human_wrt_obj", "prob_grid_canon_obj_wrt_human"):
        tol = 1e-4 * np.abs(r[k]) + 1e-30 + S * 2.0 ** -31      # cone-limited K3: terms < 2^-32 are dropped (include/coma_b200.h)
        assert (np.abs(m[k].astype(np.float64) - r[k]) <= tol).all(), k
    assert r["significant_contact_count"].sum() > 0
    for typ in ("human", "obj"):
        a, ia = get_aggregated_contact(mc, typ, 0.03)      # ratio 0.03 x 32 samples: every pair hit at least once is significant
        b, ib = ref.get_aggregated_contact(rc, typ, 0.03)
        np.testing.assert_array_equal(ia, ib)
        assert len(ib) > 0 and np.asarray(b).max() > 0
        np.testing.assert_allclose(a, b, rtol=1e-4, atol=1e-12)
        # entropy read-out (src/coma/extract_coma.py:460: n_bin = 1e6): round-half flips move a score by ~1e-6 each
        np.testing.assert_allclose(np.asarray(get_nonphysical_score(mc, typ)), np.asarray(ref.coma.get_nonphysical_score(rc, typ)), atol=2e-5)
