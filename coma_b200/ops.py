"""Typed torch-level wrappers over the C ABI (include/coma_b200.h). Tensors must live on a CUDA device;
torch is only the owner of device memory and streams here."""
import torch

from . import _lib
from ._lib import _host3, _ptr, _stream, call


def _chk(t, dtype, name):
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    return t


def nearest_vertex(pts, verts):
    """utils/coma.py:88-91. pts [N,3] f64, verts [V,3] f64 (cuda) -> int64 [N]."""
    pts, verts = _chk(pts, torch.float64, "pts").contiguous(), _chk(verts, torch.float64, "verts").contiguous()
    out = torch.empty(pts.shape[0], dtype=torch.int64, device=pts.device)
    with torch.cuda.device(pts.device):
        call("coma_nearest_vertex_f64", _ptr(pts), pts.shape[0], _ptr(verts), verts.shape[0], _ptr(out), _stream())
    return out


SUM_ORDERS = {"cpu": 0, "cuda": 1}


def pair_accumulate(hv, ov, thres, grid_size, count, nom, sum_order="cpu"):
    """utils/coma.py:284-291 over S samples. hv [S,H,3], ov [S,O,3] f32; count, nom [H,O] f32 updated in place.
    sum_order: which torch device's association of `sum(square(h - o), -1)` the squared distance follows ("cpu": (x2+y2)+z2,
    "cuda": (x2+z2)+y2) — decides `count` for pairs within one ulp of the threshold (include/coma_b200.h)."""
    S, H, _ = hv.shape
    O = ov.shape[1]
    assert ov.shape[0] == S and count.shape == (H, O) and nom.shape == (H, O)
    for t, n in ((hv, "hv"), (ov, "ov"), (count, "count"), (nom, "nom")):
        _chk(t, torch.float32, n)
    with torch.cuda.device(hv.device):
        call("coma_pair_accumulate_order_f32", _ptr(hv), _ptr(ov), S, H, O, float(thres), float(grid_size), SUM_ORDERS[sum_order],
             _ptr(count), _ptr(nom), _stream())


def bin_patches(grid_host, device=None):
    """Host-side grouping of the bin centres [N,3] (numpy fp64, N <= 256) into compact 32-bin patches for the cone-limited K3
    kernel -> int32 tensor [32 * ceil(N/32)] (on `device` if given), -1 = empty slot. No GPU work."""
    import ctypes
    import numpy as np
    g = np.ascontiguousarray(np.asarray(grid_host, dtype=np.float64).reshape(-1, 3))
    n = g.shape[0]
    perm = np.empty(32 * ((n + 31) // 32), dtype=np.int32)
    call("coma_orient_bin_patches", g.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), n, perm.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
    t = torch.from_numpy(perm)
    return t.to(device) if device is not None else t


def orient_accumulate(hn, on, grid, sigma, eps, p, sub_p, PH, PO, bin_perm=None, drop_bits=0, sum_order="cpu"):
    """utils/coma.py:295-323 over S samples. hn [S,H,3], on [S,O,3] f32; grid [N,3] f64; PH, PO [H,O,N] f32 in place.
    drop_bits = 0: dense evaluation of all N bins. drop_bits = b >= 24: cone-limited kernel — scores below 2^-b are not evaluated
    (absolute error <= S * 2^-b per bin, include/coma_b200.h); bin_perm = `bin_patches(grid)` groups the bins into compact patches.
    sum_order: "cpu" / "cuda" association of the reference's 3-term sums in the canonicalisation (see pair_accumulate)."""
    S, H, _ = hn.shape
    O, N = on.shape[1], grid.shape[0]
    assert on.shape[0] == S and PH.shape == (H, O, N) and PO.shape == (H, O, N)
    for t, n in ((hn, "hn"), (on, "on"), (PH, "PH"), (PO, "PO")):
        _chk(t, torch.float32, n)
    _chk(grid, torch.float64, "grid")
    with torch.cuda.device(hn.device):
        if bin_perm is not None:
            _chk(bin_perm, torch.int32, "bin_perm")
            assert bin_perm.numel() == 32 * ((N + 31) // 32)
        # scratch for the once-per-(sample, vertex) normalisation of the normals (cached per device; the cone kernel only)
        ws = _orient_workspace(hn.device, 3 * S * (H + O)) if drop_bits else None
        call("coma_orient_accumulate_cone_ws_f32", _ptr(hn), _ptr(on), S, H, O, _ptr(grid), N, float(sigma), float(eps), _host3(p),
             _host3(sub_p), _ptr(bin_perm), int(drop_bits), SUM_ORDERS[sum_order], _ptr(PH), _ptr(PO), _ptr(ws), _stream())


_ORIENT_WS = {}


def _orient_workspace(device, n):
    ws = _ORIENT_WS.get(device)
    if ws is None or ws.numel() < n:
        ws = _ORIENT_WS[device] = torch.empty(max(n, 1 << 20), dtype=torch.float32, device=device)
    return ws


def canonicalize(a, b, p, sub_p, eps, sum_order="cpu"):
    """utils/coma.py:123-172 -> [A,B,3] f32."""
    a, b = _chk(a, torch.float32, "a").contiguous(), _chk(b, torch.float32, "b").contiguous()
    out = torch.empty((a.shape[0], b.shape[0], 3), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        call("coma_canonicalize_order_f32", _ptr(a), a.shape[0], _ptr(b), b.shape[0], _host3(p), _host3(sub_p), float(eps),
             SUM_ORDERS[sum_order], _ptr(out), _stream())
    return out


def occupancy_accumulate(hvc, centers, thr, grids):
    """utils/coma_occupancy.py:289-295 over S samples. hvc [S,H,3] f32; centers [3,Sg] f64; grids [H,Sg,Sg,Sg] f32."""
    S, H, _ = hvc.shape
    Sg = centers.shape[1]
    assert grids.shape == (H, Sg, Sg, Sg)
    _chk(hvc, torch.float32, "hvc"), _chk(centers, torch.float64, "centers"), _chk(grids, torch.float32, "grids")
    with torch.cuda.device(hvc.device):
        call("coma_occupancy_accumulate", _ptr(hvc), S, H, _ptr(centers), Sg, float(thr), _ptr(grids), _stream())


def normalize_contact_readout(P, eps, w=None, nom=None, denom=None):
    """utils/coma.py:328-330 (in place) fused with :342-356. P [H,O,N] f32 -> contact map [H,O] (or None if w is None)."""
    H, O, N = P.shape
    _chk(P, torch.float32, "P")
    cmap = torch.empty((H, O), dtype=torch.float32, device=P.device) if w is not None else None
    with torch.cuda.device(P.device):
        call("coma_normalize_contact_readout_f32", _ptr(P), H * O, N, float(eps), _ptr(w), _ptr(nom), _ptr(denom), _ptr(cmap),
             _stream())
    return cmap


def significant_pairs(count, num):
    """utils/coma.py:376-377,:407,:421 -> (sig [H,O] bool, any_o [H] bool, any_h [O] bool)."""
    H, O = count.shape
    _chk(count, torch.float32, "count")
    sig = torch.empty((H, O), dtype=torch.uint8, device=count.device)
    any_o = torch.empty(H, dtype=torch.uint8, device=count.device)
    any_h = torch.empty(O, dtype=torch.uint8, device=count.device)
    with torch.cuda.device(count.device):
        call("coma_significant_pairs", _ptr(count), H, O, float(num), _ptr(sig), _ptr(any_o), _ptr(any_h), _stream())
    return sig.view(torch.bool), any_o.view(torch.bool), any_h.view(torch.bool)


def masked_max(cmap, mask, axis):
    """utils/coma.py:412-413 (axis=1, mask over O) / :426-427 (axis=0, mask over H); zeros if the mask is empty."""
    H, O = cmap.shape
    _chk(cmap, torch.float32, "cmap")
    mask = mask.view(torch.uint8) if mask.dtype == torch.bool else mask
    out = torch.empty(H if axis == 1 else O, dtype=torch.float32, device=cmap.device)
    with torch.cuda.device(cmap.device):
        call("coma_masked_max_f32", _ptr(cmap.contiguous()), H, O, _ptr(mask.contiguous()), int(axis), _ptr(out), _stream())
    return out


def entropy_readout(P, n_bin, weights=None):
    """utils/coma.py:455-463 on a normalised grid (weights [N] f32: the `_v2` form, :529-579). P [H,O,N] f32 -> [H,O] f32."""
    H, O, N = P.shape
    out = torch.empty((H, O), dtype=torch.float32, device=P.device)
    with torch.cuda.device(P.device):
        if weights is None:
            call("coma_entropy_readout_f32", _ptr(P), H * O, N, float(n_bin), _ptr(out), _stream())
        else:
            weights = _chk(weights, torch.float32, "weights").contiguous()
            call("coma_entropy_readout_weighted_f32", _ptr(P), H * O, N, float(n_bin), _ptr(weights), float(weights.double().sum().item()),
                 _ptr(out), _stream())
    return out


def occupancy_readout(grids, sel_idx=None):
    """utils/coma_occupancy.py:297-312: normalises `grids` [H,...] in place, returns the max field over `sel_idx`."""
    H = grids.shape[0]
    V = grids[0].numel()
    field = torch.empty(grids.shape[1:], dtype=torch.float32, device=grids.device)
    nsel = 0 if sel_idx is None else sel_idx.numel()
    if sel_idx is not None and nsel == 0:
        # an EMPTY selection: a 0-element tensor has a NULL data pointer, which the C ABI reads as "no selection, use every
        # vertex" — pass a non-null dummy so the kernel really selects nothing (field = 0, grids still normalised in place)
        sel_idx = torch.zeros(1, dtype=torch.int64, device=grids.device)
    with torch.cuda.device(grids.device):
        call("coma_occupancy_readout_f32", _ptr(grids), H, V, _ptr(sel_idx), nsel, _ptr(field), _stream())
    return field


def gemm_f16(a, w, bias=None, residual=None, act=0, out_dtype=torch.float16, out=None):
    """G1: out = act(a @ w.T + bias + residual) on the tcgen05 tensor cores. a [M,K] f16, w [N,K] f16 (K contiguous),
    bias [N] f32, residual [M,N] f16; out f16 or f32 [M,N]."""
    assert a.dtype == torch.float16 and w.dtype == torch.float16 and a.dim() == 2 and w.dim() == 2
    assert a.stride(1) == 1 and w.stride(1) == 1 and a.shape[1] == w.shape[1]
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    assert out.shape == (M, N) and out.stride(1) == 1 and out.dtype in (torch.float16, torch.float32)
    if residual is not None:
        assert residual.dtype == torch.float16 and residual.shape == (M, N) and residual.stride(1) == 1 and residual.stride(0) == out.stride(0)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N and bias.is_contiguous()
    o16 = out.data_ptr() if out.dtype == torch.float16 else None
    o32 = out.data_ptr() if out.dtype == torch.float32 else None
    with torch.cuda.device(a.device):
        call("coma_gemm_f16_tn", a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), M, N, K, _ptr(bias),
             None if residual is None else residual.data_ptr(), int(act), o16, o32, out.stride(0), _stream())
    return out
