"""oracle/oracle.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy/ctypes front-end of the CPU oracle (oracle/coma_oracle.c) plus numpy restatements of the reference's read-outs.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import this module;
nothing under coma_b200/ does.  Parity status: PINNED to reference-generated vectors (tests/golden/, see
tests/test_oracle_golden.py).

Reference lines followed (relative to the reference root):
  fibonacci_sphere            utils/coma.py:18-26
  voxel_centers               utils/coma_occupancy.py:160-171
  to_f32                      utils/misc.py:37-54
  normalize_normals           utils/coma.py:328-330
  contact_map                 utils/coma.py:333-366
  significant_pairs           utils/coma.py:369-382
  aggregate_contact           utils/coma.py:385-438, 614-641
  entropy_score               utils/coma.py:441-476
  occupancy_field             utils/coma_occupancy.py:297-312
  nearest_distance / chamfer_distance / minimum_distance
                              src/application/optimize.py:155-165, src/generation/optimize_depth.py:29-44 (pinned against the
                              reference's own torch.cdist expressions in tests/test_geometry.py)
"""
import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)
_i64p = ctypes.POINTER(ctypes.c_int64)
_i64 = ctypes.c_int64


def build(force=False):
    so = os.path.join(_HERE, "_build", "liboracle.so")
    src = os.path.join(_HERE, "coma_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.check_call(["make", "-C", _HERE, "-B", "_build/liboracle.so"], env=env,
                              stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        L.oracle_num_threads.restype = ctypes.c_int
        L.oracle_set_num_threads.argtypes = [ctypes.c_int]
        L.oracle_set_sum_order.argtypes = [ctypes.c_int]
        L.oracle_nearest_vertex_f64.argtypes = [_f64p, _i64, _f64p, _i64, _i64p]
        L.oracle_pair_accumulate_f32.argtypes = [_f32p, _f32p, _i64, _i64, _i64, ctypes.c_float, ctypes.c_float, _f32p, _f32p]
        L.oracle_pair_accumulate_order_f32.argtypes = [_f32p, _f32p, _i64, _i64, _i64, ctypes.c_float, ctypes.c_float, ctypes.c_int, _f32p, _f32p]
        L.oracle_canonicalize_f32.argtypes = [_f32p, _i64, _f32p, _i64, _f32p, _f32p, ctypes.c_float, _f32p]
        L.oracle_orient_accumulate.argtypes = [_f32p, _f32p, _i64, _i64, _i64, _f64p, _i64, ctypes.c_double, ctypes.c_double,
                                               _f32p, _f32p, _f32p, _f32p]
        L.oracle_occupancy_accumulate.argtypes = [_f32p, _i64, _i64, _f64p, _i64, ctypes.c_double, _f32p]
        _LIB = L
    return _LIB


def num_threads():
    return lib().oracle_num_threads()


def set_num_threads(n):
    lib().oracle_set_num_threads(int(n))


def _c(a, dt):
    a = np.ascontiguousarray(a, dtype=dt)
    return a


def _p(a, t):
    return a.ctypes.data_as(t)


# ----------------------------------------------------------------------------------------------- host tables
def fibonacci_sphere(n):
    """utils/coma.py:18-26 -> float64 [n,3] (ComA.canon_normal_grid before any load())."""
    indices = np.arange(0, n, dtype=float) + 0.5
    phi = np.arccos(1 - 2 * indices / n)
    theta = np.pi * (1 + 5**0.5) * indices
    return np.stack([np.cos(theta) * np.sin(phi), np.sin(theta) * np.sin(phi), np.cos(phi)], axis=-1)


def voxel_centers(Sg, gridsize=2.4):
    """Per-axis voxel centres of load_voxelgrid (utils/coma_occupancy.py:160-171): the middle term is an fp32 product
    (python float x float32 array stays float32 in numpy >= 2) promoted to fp64 by `start_point`."""
    voxel = gridsize / Sg
    start = -gridsize / 2.0
    idx = np.arange(Sg).astype(np.float32)
    c = np.float64(start) + (voxel * idx).astype(np.float64) + voxel / 2
    return np.stack([c, c, c]), voxel


def to_f32(x):
    """utils/misc.py:37-54: every float array becomes fp32 before it reaches the accumulation code."""
    return np.ascontiguousarray(np.asarray(x), dtype=np.float32)


# ----------------------------------------------------------------------------------------------- accumulation
def vertex_normals(verts, faces, eps=None):
    """Sample ingest, reference utils/coma.py:665-686: open3d `TriangleMesh.compute_vertex_normals()` — the sum
    over incident faces of the UN-normalised face normals (v1-v0)x(v2-v0), normalised, (0,0,1) where the sum has no finite
    direction [open3d 0.17 TriangleMesh.cpp ComputeVertexNormals; open3d is not installed here: PARITY UNPINNED against
    open3d itself] — optionally followed by `normalize_vectors_np(., eps)` (utils/transformations.py:8-11). fp64, [.., V, 3]."""
    v = np.asarray(verts, dtype=np.float64)
    f = np.asarray(faces, dtype=np.int64).reshape(-1, 3)
    lead = v.shape[:-2]
    v = v.reshape(-1, v.shape[-2], 3)
    out = np.empty_like(v)
    for s in range(v.shape[0]):
        fn = np.cross(v[s][f[:, 1]] - v[s][f[:, 0]], v[s][f[:, 2]] - v[s][f[:, 0]])
        vn = np.zeros_like(v[s])
        for k in range(3):
            np.add.at(vn, f[:, k], fn)
        n = np.sqrt(np.sum(np.square(vn), axis=-1, keepdims=True))
        with np.errstate(invalid="ignore", divide="ignore"):
            o = vn / np.where(n > 0, n, 1.0)
        o[(n[:, 0] == 0) | ~np.isfinite(o).all(-1)] = (0.0, 0.0, 1.0)
        if eps is not None:
            o = o / (np.sqrt(np.sum(np.square(o), axis=-1, keepdims=True)) + eps)
        out[s] = o
    return out.reshape(*lead, v.shape[-2], 3)


def nearest_vertex(pts, verts):
    pts, verts = _c(pts, np.float64), _c(verts, np.float64)
    out = np.empty(len(pts), dtype=np.int64)
    lib().oracle_nearest_vertex_f64(_p(pts, _f64p), len(pts), _p(verts, _f64p), len(verts), _p(out, _i64p))
    return out


def pair_accumulate(hv, ov, thres, grid_size, count=None, nom=None, sum_order="cpu"):
    """utils/coma.py:284-291. sum_order: the association of torch.sum(torch.square(h - o), -1) — "cpu" ((xx+yy)+zz, also numpy) or
    "cuda" ((xx+zz)+yy, ATen's CUDA reduction as measured on B200; pinned by tests/golden/cuda_*.npz)."""
    hv, ov = to_f32(hv), to_f32(ov)
    S, H, _ = hv.shape
    O = ov.shape[1]
    count = np.zeros((H, O), np.float32) if count is None else count
    nom = np.zeros((H, O), np.float32) if nom is None else nom
    lib().oracle_pair_accumulate_order_f32(_p(hv, _f32p), _p(ov, _f32p), S, H, O, np.float32(thres), np.float32(grid_size),
                                           {"cpu": 0, "cuda": 1}[sum_order], _p(count, _f32p), _p(nom, _f32p))
    return count, nom


def canonicalize(a, b, p=(0, 0, 1), sub_p=(0, 1, 0), eps=1e-8, sum_order="cpu"):
    lib().oracle_set_sum_order({"cpu": 0, "cuda": 1}[sum_order])
    a, b = to_f32(a), to_f32(b)
    p, sp = to_f32(p), to_f32(sub_p)
    out = np.empty((len(a), len(b), 3), np.float32)
    lib().oracle_canonicalize_f32(_p(a, _f32p), len(a), _p(b, _f32p), len(b), _p(p, _f32p), _p(sp, _f32p), np.float32(eps),
                                  _p(out, _f32p))
    return out


def orient_accumulate(hn, on, grid, sigma, eps, p=(0, 0, 1), sub_p=(0, 1, 0), PH=None, PO=None, sum_order="cpu"):
    hn, on = to_f32(hn), to_f32(on)
    grid = _c(grid, np.float64)
    S, H, _ = hn.shape
    O, N = on.shape[1], len(grid)
    PH = np.zeros((H, O, N), np.float32) if PH is None else PH
    PO = np.zeros((H, O, N), np.float32) if PO is None else PO
    p, sp = to_f32(p), to_f32(sub_p)
    lib().oracle_set_sum_order({"cpu": 0, "cuda": 1}[sum_order])
    lib().oracle_orient_accumulate(_p(hn, _f32p), _p(on, _f32p), S, H, O, _p(grid, _f64p), N, float(sigma), float(eps),
                                   _p(p, _f32p), _p(sp, _f32p), _p(PH, _f32p), _p(PO, _f32p))
    return PH, PO


def occupancy_accumulate(human_verts, obj_verts, Sg, scale_tolerance, grids=None):
    """human_verts [S,H,3], obj_verts [S,O,3] in the caller's dtype (fp64 in the real pipeline)."""
    hv, ov = np.asarray(human_verts), np.asarray(obj_verts)
    hvc = to_f32(hv - ov[:, 0:1, :])                       # :287-288 host subtraction in the input dtype, then fp32
    S, H, _ = hvc.shape
    centers, voxel = voxel_centers(Sg)
    centers = _c(centers, np.float64)
    grids = np.zeros((H, Sg, Sg, Sg), np.float32) if grids is None else grids
    lib().oracle_occupancy_accumulate(_p(hvc, _f32p), S, H, _p(centers, _f64p), Sg, float(voxel * scale_tolerance),
                                      _p(grids, _f32p))
    return grids


# ----------------------------------------------------------------------------------------------- point-set distances (8f-3)
def nearest_distance(a, b, block=2048):
    """min_j ||a_i - b_j|| and the first arg-min, fp32 with separately rounded ops in the order (dx^2+dy^2)+dz^2 — the row minima of
    torch.cdist(a, b, compute_mode="donot_use_mm_for_euclid_dist") that `chamfer_distance` (src/application/optimize.py:155-165) and
    `minimum_distance` (src/generation/optimize_depth.py:29-44) reduce. Blocked over a to bound memory."""
    a, b = to_f32(a), to_f32(b)
    dist = np.empty(len(a), np.float32)
    idx = np.empty(len(a), np.int32)
    for i0 in range(0, len(a), block):
        d = a[i0:i0 + block, None, :] - b[None, :, :]
        sq = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
        j = np.argmin(sq, axis=1)
        idx[i0:i0 + block] = j
        dist[i0:i0 + block] = np.sqrt(sq[np.arange(len(j)), j])
    return dist, idx


def chamfer_distance(a, b):
    """src/application/optimize.py:155-165."""
    return np.float32(nearest_distance(a, b)[0].mean(dtype=np.float32) + nearest_distance(b, a)[0].mean(dtype=np.float32))


def minimum_distance(a, b, num_vertices=100):
    """src/generation/optimize_depth.py:29-44."""
    return np.float32(np.sort(nearest_distance(a, b)[0])[:num_vertices].mean(dtype=np.float32))


# ----------------------------------------------------------------------------------------------- read-outs
def normalize_normals(P, eps):
    """utils/coma.py:328-330 (in place in the reference; returns a new fp32 array here)."""
    s = P.sum(axis=-1, keepdims=True, dtype=np.float32) + np.float32(eps)
    return (P / s).astype(np.float32)


def contact_map(Pn, grid, nom, denom, p=(0, 0, 1)):
    """utils/coma.py:342-356 on an already-normalised grid; fp64 when `grid` is fp64 (type promotion)."""
    p = to_f32(p)
    dots = (p[None, :] * grid).sum(-1)
    w = (1.0 - dots) / 2.0
    return (Pn * w[None, None, :]).sum(-1) * (nom / denom)


def significant_pairs(count, ratio, used_count):
    return count >= ratio * used_count


def aggregate_contact(cmap, sig, which):
    """utils/coma.py:398-427 + 633-639 -> (fp32 map, int64 indices)."""
    H, O = cmap.shape
    if which == "human":
        sel = sig.any(0)
        agg = cmap[:, sel].max(-1) if sel.any() else np.zeros(H)
    else:
        sel = sig.any(1)
        agg = cmap[sel, :].max(0) if sel.any() else np.zeros(O)
    return agg.astype(np.float32), np.argwhere(sel)[:, 0]


def entropy_score(Pn, n_bin=1e6):
    """utils/coma.py:455-463 on an already-normalised fp32 grid."""
    q = (np.round(Pn * np.float32(n_bin)) / np.float32(n_bin)).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = np.where(q == 0, np.float32(0), q * np.log(q)).astype(np.float32)
    return (t.sum(-1, dtype=np.float32) / np.float32(math.log(n_bin)) + np.float32(1.0)).astype(np.float32)


def occupancy_field(grids):
    """utils/coma_occupancy.py:297-312: per-vertex normalisation (NaN where a vertex never hit), max over vertices."""
    H = grids.shape[0]
    flat = grids.reshape(H, -1)
    with np.errstate(divide="ignore", invalid="ignore"):
        norm = flat / flat.sum(-1, keepdims=True, dtype=np.float32)
    norm = norm.reshape(grids.shape)
    # torch.max propagates NaN
    return np.where(np.isnan(norm).any(0), np.float32(np.nan), norm.max(0)).astype(np.float32), norm
