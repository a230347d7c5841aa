"""Host-side logic that needs no GPU: drop-in import paths, pickle layout, cache bookkeeping, voxel-grid tables,
fp32 boundary semantics, sharding rules, and the loud failure when the CUDA path is unavailable."""
import pickle

import numpy as np
import pytest
import torch


def _coma(H=6, O=4, N=16):
    from utils.coma import ComA
    return ComA(human_res=H, obj_res=O, normal_res=N, spatial_res=0,
                proximity_settings=dict(spatial_grid_size=0.07, spatial_grid_thres=0.03),
                normal_gaussian_sigma=0.25, eps=1e-10, device="cpu")


def test_dropin_import_paths():
    import utils.coma as uc
    import utils.coma_occupancy as uo
    import utils.misc as um
    for name in ("ComA", "get_aggregated_contact", "negative_exp", "get_uniform_points_on_sphere"):
        assert hasattr(uc, name)
    assert hasattr(uo, "ComA_Occupancy") and hasattr(uo, "load_voxelgrid")
    assert hasattr(um, "to_np_torch_recursive") and hasattr(um, "get_3d_indexgrid_ijk")


def test_export_layout_matches_reference(golden_dir):
    """Key list, dtypes and shapes of ComA.export (measured on the reference)."""
    c = _coma()
    e = c.export()
    assert list(e.keys()) == [
        "device", "human_res", "obj_res", "normal_res", "spatial_res", "canon_normal_grid",
        "prob_grid_canon_human_wrt_obj", "prob_grid_canon_obj_wrt_human", "contact_dist_expectation_grid_nom",
        "contact_dist_expectation_grid_denom", "significant_contact_count", "proximity_settings", "contact_dist_func",
        "cross_contact_scores_nom", "cross_contact_scores_denom", "cache_count", "used_count", "principle_vec",
        "sub_principle_vec", "rel_dist_method", "normal_gaussian_sigma", "eps"]
    assert e["canon_normal_grid"].dtype == np.float32 and e["canon_normal_grid"].shape == (16, 3)
    assert e["prob_grid_canon_human_wrt_obj"].shape == (6, 4, 16)
    for k in ("contact_dist_expectation_grid_nom", "contact_dist_expectation_grid_denom", "significant_contact_count",
              "cross_contact_scores_nom", "cross_contact_scores_denom"):
        assert e[k].dtype == np.float32 and e[k].shape == (6, 4)
    assert e["principle_vec"].tolist() == [0, 0, 1] and e["sub_principle_vec"].tolist() == [0, 1, 0]
    # the reference's grid, bit for bit
    g = np.load(f"{golden_dir}/contact_small.npz")
    from utils.coma import ComA
    big = ComA(2, 2, int(g["N"]), 0, proximity_settings=dict(spatial_grid_size=1, spatial_grid_thres=1), device="cpu")
    assert big.canon_normal_grid.dtype == torch.float64
    np.testing.assert_array_equal(big.export()["canon_normal_grid"], g["canon_normal_grid"])


def test_pickle_roundtrip_and_negative_exp_path(tmp_path):
    c = _coma()
    c.used_count = 7
    c.significant_contact_count += 3
    p = tmp_path / "c.pickle"
    c.export(save_pth=str(p))
    raw = pickle.load(open(p, "rb"))
    fn = raw["contact_dist_func"]
    assert fn.func.__module__ == "utils.coma" and fn.func.__name__ == "negative_exp"   # reference pickles load here and vice versa
    assert fn.keywords == dict(spatial_grid_size=0.07, spatial_grid_thres=0.03)
    x = torch.tensor([0.0, 0.07])
    torch.testing.assert_close(fn(x), torch.exp(-x / 0.07))
    d = _coma()
    d.load(str(p))
    assert d.used_count == 7 and d.significant_contact_count.dtype == torch.float32
    assert d.canon_normal_grid.dtype == torch.float32          # reference behaviour after a load() round trip
    assert torch.equal(d.significant_contact_count, torch.full((6, 4), 3.0))


def test_cache_bookkeeping_and_borrowed_arrays():
    c = _coma()
    arrs = dict(human_verts=np.zeros((6, 3)), human_normals=np.ones((6, 3)), obj_verts=np.zeros((4, 3)), obj_normals=np.ones((4, 3)))
    c.register_sample_to_cache(**arrs)
    c.register_sample_to_cache(**arrs)
    assert c.cache_count == 2 and list(c.cache) == ["00000", "00001"]
    assert c.cache["00000"]["human_verts"] is arrs["human_verts"]     # borrowed, not copied


def test_no_cpu_fallback():
    c = _coma()
    c.register_sample_to_cache(human_verts=np.zeros((6, 3)), human_normals=np.ones((6, 3)), obj_verts=np.zeros((4, 3)), obj_normals=np.ones((4, 3)))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        c.aggregate_all_samples()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        c.compute_contact_map("human")
    from utils.coma_occupancy import ComA_Occupancy
    o = ComA_Occupancy(3.0, 6, 4, 0, 5, device="cpu")
    o.register_sample_to_cache(human_verts=np.zeros((6, 3)), human_normals=np.ones((6, 3)), obj_verts=np.zeros((4, 3)), obj_normals=np.ones((4, 3)))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        o.aggregate_all_samples()


def test_input_asserts_and_not_implemented():
    from utils.coma import ComA
    from utils.coma_occupancy import ComA_Occupancy
    c = _coma()
    with pytest.raises(AssertionError):
        c.assert_inputs(human_verts=np.zeros((5, 3)))
    with pytest.raises(AssertionError):
        c.assert_inputs(obj_normals=np.zeros((4, 2)))
    with pytest.raises(AssertionError):
        c.assert_inputs(contact_map_type="x")
    with pytest.raises(NotImplementedError):
        ComA(2, 2, 4, spatial_res=2, device="cpu")
    with pytest.raises(AssertionError):
        ComA(2, 2, 4, 0, rel_dist_method="nope", device="cpu")
    with pytest.raises(AssertionError):
        ComA_Occupancy(3.0, 2, 2, normal_res=4, spatial_res=4, device="cpu")


def test_occupancy_grid_tables_and_export(golden_dir):
    from utils.coma_occupancy import ComA_Occupancy, load_voxelgrid
    g = np.load(f"{golden_dir}/occupancy_small.npz")
    Sg = int(g["Sg"])
    o = ComA_Occupancy(float(g["tol"]), 5, 3, 0, Sg, device="cpu")
    e = o.export()
    np.testing.assert_array_equal(e["spatial_grid"], g["spatial_grid"])
    assert e["spatial_grid"].dtype == np.float32 and e["spatial_indexgrid"].dtype == np.int64
    assert e["spatial_indexgrid"].shape == (3, Sg, Sg, Sg) and e["spatial_indexgrid"][:, 1, 2, 3].tolist() == [1, 2, 3]
    assert e["rel_dist_thres"] == float(g["rel_dist_thres"])
    md = e["spatial_grid_metadata"]
    assert md["start_point"].dtype == np.float32 and md["voxel_size"] == float(g["voxel_size"]) and md["N_x"] == Sg
    assert o._centers.dtype == torch.float64
    grid, idx, meta = load_voxelgrid(2.4, 30)
    # the fp32-rounded middle term of the reference's expression (utils/coma_occupancy.py:171)
    exact = -1.2 + (2.4 / 30) * np.arange(30) + (2.4 / 30) / 2
    assert np.abs(grid[0, :, 0, 0] - exact).max() > 0 and np.abs(grid[0, :, 0, 0] - exact).max() < 1e-6


def test_to_np_torch_recursive_forces_fp32_int64():
    from utils.misc import to_np_torch_recursive
    d = dict(a=np.array([0.1], dtype=np.float64), b=[np.array([1], dtype=np.int32)], c=dict(t=torch.tensor([0.1], dtype=torch.float64)))
    t = to_np_torch_recursive(d, use_torch=True, device="cpu")
    assert t["a"].dtype == torch.float32 and t["b"][0].dtype == torch.int64 and t["c"]["t"].dtype == torch.float32
    assert t["a"].item() == float(np.float32(0.1))
    n = to_np_torch_recursive(t, use_torch=False, device="cpu")
    assert n["a"].dtype == np.float32 and n["b"][0].dtype == np.int64


def test_sharding_rules():
    from coma_b200 import dist
    assert sorted(sum((dist.sample_shard(37, r, 8) for r in range(8)), [])) == list(range(37))
    sl = [dist.human_slice(10475, r, 8) for r in range(8)]
    assert sl[0][0] == 0 and sl[-1][1] == 10475 and all(sl[i][1] == sl[i + 1][0] for i in range(7))
    assert max(b - a for a, b in sl) - min(b - a for a, b in sl) <= 1
    # the reference's rule: sub = n//N + 1 (src/generation/inpaint.py:272-278)
    n, N = 100, 8
    got = [dist.work_item_slice(n, i, N) for i in range(N)]
    items = list(range(n))
    sub = n // N + 1
    assert [items[a:b] for a, b in got] == [items[i * sub:(i + 1) * sub] for i in range(N)]
    assert sum(b - a for a, b in got) == n


def test_geglu_weight_interleave():
    """prep_geglu reorders the feed-forward projection rows into blocks of 32 values followed by their 32 gates (the layout
    the fused GEGLU epilogue reads from one 64-column accumulator group)."""
    import torch
    from coma_b200.inpaint import nn
    F, C = 128, 24
    w = torch.arange(2 * F * C, dtype=torch.float32).reshape(2 * F, C) / 1024
    b = torch.arange(2 * F, dtype=torch.float32)
    wi, bi = nn.prep_geglu(w, b, "cpu")
    assert wi.shape == (2 * F, C) and bi.shape == (2 * F,)
    for j in range(F // 32):
        assert torch.equal(bi[64 * j: 64 * j + 32], b[32 * j: 32 * j + 32])                  # values of features 32j..32j+31
        assert torch.equal(bi[64 * j + 32: 64 * j + 64], b[F + 32 * j: F + 32 * j + 32])      # their gates
        assert torch.equal(wi[64 * j + 5].float(), w[32 * j + 5].half().float())
        assert torch.equal(wi[64 * j + 37].float(), w[F + 32 * j + 5].half().float())
    assert nn.prep_geglu(torch.zeros((2 * 40, C)), torch.zeros(2 * 40), "cpu") is None        # 2F % 256 != 0 -> unfused path


def test_oracle_vertex_normals_matches_cli_helper():
    """The oracle restatement of open3d's vertex normals (K6 checker) agrees with the helper the CLI used before K6 existed,
    and normalize_vectors_np is applied on top when eps is given."""
    from coma_b200.cli.io import vertex_normals
    from coma_b200.misc import normalize_vectors_np
    from oracle import oracle
    rng = np.random.default_rng(0)
    v = rng.standard_normal((50, 3))
    f = rng.integers(0, 49, (120, 3))          # vertex 49 is isolated -> (0, 0, 1)
    a, b = oracle.vertex_normals(v, f), vertex_normals(v, f)
    assert np.array_equal(a, b) and a[49].tolist() == [0.0, 0.0, 1.0]
    assert np.array_equal(oracle.vertex_normals(v, f, 1e-10), normalize_vectors_np(b, eps=1e-10))
    assert oracle.vertex_normals(np.stack([v, 2 * v]), f).shape == (2, 50, 3)


def test_gemm_planner_choices():
    """The tile-width / split-K cost model (host code, no GPU needed): problems with too few 128-row tiles for 148 SMs are narrowed
    or split along K; big problems keep wide tiles and never split; no workspace -> never split."""
    import ctypes
    from coma_b200 import _lib
    lib = _lib.load()

    def plan(M, N, K, ws=8 << 20, conv_m_tiles=0):
        bn, ks = ctypes.c_int(0), ctypes.c_int(0)
        assert lib.coma_gemm_plan(M, N, K, ws, conv_m_tiles, ctypes.byref(bn), ctypes.byref(ks)) == 0
        return bn.value, ks.value

    assert plan(8192, 8192, 8192) == (256, 1)                      # plenty of tiles
    bn, ks = plan(512, 1280, 11520)                                # 8x8-level 3x3 conv: 4 x 5 tiles of 180 K-slabs
    assert ks >= 2 and bn in (128, 160, 256)                       # split along K (slab traffic caps the factor)
    assert plan(512, 1280, 11520, ws=0)[1] == 1                    # no scratch, no split
    assert plan(2048, 1280, 1280) == (160, 1)                      # 128 tiles in one wave instead of 80 wide ones
    assert plan(32768, 320, 320) == (160, 1)                       # N = 320: two exact 160-wide tiles
    assert plan(512, 1280, 1280)[0] == 64                          # short K: narrow tiles rather than a split
    assert plan(1 << 20, 128, 1152) == (128, 1)
    for M, N, K in [(77, 320, 768), (616, 640, 768), (32768, 2560, 320), (8192, 640, 2560)]:
        bn, ks = plan(M, N, K)
        assert bn in (64, 128, 160, 256) and 1 <= ks <= 16 and ks * M * N <= 8 << 20 or ks == 1


def test_ingest_corner_list_reproduces_numpy_accumulation_order():
    """K6's host side: gathering each vertex's faces in the corner-list order and adding sequentially gives bit-for-bit what
    np.add.at(vn, f[:, k], fn), k = 0, 1, 2 gives (the oracle's accumulation) — including vertices without faces."""
    from coma_b200.ingest import build_corner_list
    rng = np.random.default_rng(1)
    V, F = 60, 200
    faces = rng.integers(0, V - 1, (F, 3))                 # vertex V-1 has no faces
    fn = rng.standard_normal((F, 3)) * np.exp(rng.uniform(-8, 8, (F, 1)))   # wide dynamic range: order matters in fp64
    f32, off, cf = build_corner_list(faces, V)
    assert off[0] == 0 and off[-1] == 3 * F and f32.dtype == np.int32 and off[V] == off[V - 1]
    ref = np.zeros((V, 3))
    for k in range(3):
        np.add.at(ref, faces[:, k], fn)
    mine = np.zeros((V, 3))
    for v in range(V):
        for i in range(off[v], off[v + 1]):
            mine[v] = mine[v] + fn[cf[i]]
    assert np.array_equal(mine, ref)
    # every corner appears exactly once
    assert sorted(cf.tolist()) == sorted(np.repeat(np.arange(F), 3).tolist())


def test_host_stage_rows_helper():
    """`coma_host_stage_rows_f64_f32` / `coma_host_rows_equal_f64` (host code of the library, no GPU): same values as the reference's
    `(human_verts - obj_verts[0]).astype(float32)` (utils/coma_occupancy.py:287-288) and `to_np_torch_recursive`'s fp64 -> fp32 rounding
    (utils/misc.py:47-54); first-mismatch reporting; None (= numpy fallback) for arrays that are not plain float64."""
    from coma_b200.staging import rows_equal_f64, stage_rows_f64
    rng = np.random.default_rng(5)
    H, n = 257, 9
    hv = [rng.normal(size=(H, 3)) * 10.0 ** rng.integers(-3, 3) for _ in range(n)]
    ov = [rng.normal(size=(4, 3)) for _ in range(n)]
    out = np.full((n + 2, 100, 3), -7.0, np.float32)
    assert stage_rows_f64(hv, out[:n], row0=50, sub=ov, equal_to=ov[0][0]) == 1          # sample 1 has another object
    want = np.stack([(h[50:150] - o[0][None]).astype(np.float32) for h, o in zip(hv, ov)])
    assert np.array_equal(out[:n], want) and (out[n:] == -7.0).all()
    same = [ov[0]] * n
    assert stage_rows_f64(hv, out[:n], row0=0, sub=same, equal_to=ov[0][0]) == -1
    assert stage_rows_f64(hv, out[:n], row0=157, sub=None) == -1
    assert np.array_equal(out[:n], np.stack([h[157:].astype(np.float32) for h in hv]))
    assert stage_rows_f64(hv, out[:n], row0=200) is None                                  # not enough rows
    assert stage_rows_f64([h.astype(np.float32) for h in hv], out[:n]) is None            # wrong dtype -> caller falls back to numpy
    assert stage_rows_f64([h[:, ::-1] for h in hv], out[:n]) is None                      # not contiguous
    assert rows_equal_f64(same, ov[0][0]) is True and rows_equal_f64(ov, ov[0][0]) is False
    assert rows_equal_f64([o.astype(np.float32) for o in ov], ov[0][0]) is None


def test_host_stage_rows_addresses():
    """The data pointers handed to the host helper come from the buffer protocol (`_addresses`): read-only arrays (no writable buffer
    export), offset views, and a run of identical objects (one shared object mesh — looked up once) must all stage the right values."""
    from coma_b200.staging import _addresses, rows_equal_f64, stage_rows_f64
    rng = np.random.default_rng(11)
    H, n = 64, 6
    hv = [rng.normal(size=(H + 5, 3)) for _ in range(n)]
    ro = hv[2]
    ro.flags.writeable = False                                                            # e.g. arrays that came out of np.load(mmap_mode="r")
    views = [h[5:] for h in hv]                                                           # contiguous views that do not start at the base pointer
    assert _addresses(views) == [v.ctypes.data for v in views]
    assert _addresses([hv[0], hv[0], hv[1], hv[0]]) == [hv[0].ctypes.data, hv[0].ctypes.data, hv[1].ctypes.data, hv[0].ctypes.data]
    obj = rng.normal(size=(3, 3))
    out = np.zeros((n, H, 3), np.float32)
    assert stage_rows_f64(views, out, row0=0, sub=[obj] * n, equal_to=obj[0]) == -1
    assert np.array_equal(out, np.stack([(v - obj[0][None]).astype(np.float32) for v in views]))
    assert rows_equal_f64([obj[1:]] * 3 + [obj[1:].copy()], obj[1]) is True                # offset view of the shared object, then an equal copy


def test_occupancy_chunk_staging_matches_numpy_path():
    """ComA_Occupancy._stage_chunk (one library call per chunk) writes exactly what the per-sample numpy path writes, and declines
    (-> numpy path, which asserts / np.allclose's like the reference) when the object moves or an input is not float64."""
    from coma_b200.coma_occupancy import ComA_Occupancy
    rng = np.random.default_rng(11)
    H, O, n = 64, 5, 6
    ov, on = rng.normal(size=(O, 3)), rng.normal(size=(O, 3))
    samples = [dict(human_verts=rng.normal(size=(H, 3)), human_normals=rng.normal(size=(H, 3)), obj_verts=ov.copy(), obj_normals=on.copy()) for _ in range(n)]
    occ = ComA_Occupancy.__new__(ComA_Occupancy)
    occ.human_res, occ._human_slice, occ.debug_obj_vert, occ.debug_obj_normal = H, (8, 40), None, None
    fast = np.zeros((n, 32, 3), np.float32)
    assert occ._stage_chunk(samples, fast, all_rows=False)
    slow = np.zeros_like(fast)
    for j, s in enumerate(samples):
        occ._canonical_human_verts(s, out=slow[j])
    assert np.array_equal(fast, slow)
    full = np.zeros((n, H, 3), np.float32)
    assert occ._stage_chunk(samples, full, all_rows=True) and np.array_equal(full[:, 8:40], slow)
    moved = [dict(s) for s in samples]
    moved[3]["obj_verts"] = ov + 1e-12
    assert not occ._stage_chunk(moved, fast, all_rows=False)
    f32 = [dict(s, human_verts=s["human_verts"].astype(np.float32)) for s in samples]
    assert not occ._stage_chunk(f32, fast, all_rows=False)


@pytest.mark.parametrize("Sg,tol", [(7, 1.2), (30, 3.0), (64, 2.5), (128, 3.0), (2040, 3.0), (2040, 0.6)])
def test_occupancy_candidate_box_contains_every_hit(Sg, tol):
    """K4 (coma_b200/csrc/occupancy.cu: axis_range) tests only the voxels of an index box derived from a LINEAR estimate of the centre
    positions: [ceil(x_lo - 1/64), floor(x_hi + 1/64)], x = (v -+ thr - c[0]) * (Sg - 1) / (c[Sg-1] - c[0]). The reference's centres
    are not exactly uniform (fp32 middle term, utils/coma_occupancy.py:171), so the claim "the box contains every voxel the reference
    would count" is checked here against the exact per-axis hit set, for random vertices and for vertices planted at centre +- thr
    (+- one fp64 ulp, then rounded to fp32 like the staged input), from the smallest grid to the largest the kernel accepts."""
    from oracle import oracle
    centers, voxel = oracle.voxel_centers(Sg)
    c = np.asarray(centers[0], np.float64)
    thr = voxel * tol
    T = thr * thr                                       # smallest double whose sqrt is >= thr (squared_threshold in occupancy.cu)
    while np.sqrt(T) >= thr:
        T = np.nextafter(T, 0.0)
    while np.sqrt(T) < thr:
        T = np.nextafter(T, np.inf)
    rng = np.random.default_rng(Sg)
    pick = c if Sg <= 512 else c[rng.integers(0, Sg, 512)]
    v = np.concatenate([rng.uniform(-1.4, 1.4, 4000), pick + thr, pick - thr, np.nextafter(pick + thr, 9.0),
                        np.nextafter(pick - thr, -9.0), pick, pick + voxel / 2]).astype(np.float32).astype(np.float64)
    inv = (Sg - 1) / (c[-1] - c[0])
    lo = np.maximum(np.ceil((v - thr - c[0]) * inv - 1.0 / 64.0), 0).astype(np.int64)
    hi = np.minimum(np.floor((v + thr - c[0]) * inv + 1.0 / 64.0), Sg - 1).astype(np.int64)
    idx = np.arange(Sg)[None, :]
    for a in range(0, len(v), 1024):
        d = c[None, :] - v[a:a + 1024, None]
        hit = (d * d) < T                               # necessary for a 3-D hit: the other two squares only add
        inbox = (idx >= lo[a:a + 1024, None]) & (idx <= hi[a:a + 1024, None])
        assert not (hit & ~inbox).any()
    assert (hi - lo + 1).max() <= int(np.ceil(2 * tol)) + 2          # and it is tight: at most one spare voxel per side
