"""ComA_Occupancy — drop-in mirror of the reference's `utils/coma_occupancy.py` (:160-343) on B200.

Per-human-vertex voxel-hit counting relative to object vertex 0. Same constructor, methods, asserts and pickle layout;
the dense [H, Sg^3] fp64 distance test of the reference is replaced by the K4 scatter kernel (bit-exact counts) and the
read-out by the K5c kernels.  Multi-GPU: shard the HUMAN-VERTEX axis (`human_slice`), every rank sees all samples and
owns H/world full grids; the only collective is a MAX all-reduce of the [Sg^3] field.
"""
import pickle

import numpy as np
import torch

from . import dist as cdist
from . import ops
from .misc import get_3d_indexgrid_ijk, to_np_torch_recursive
from .staging import BatchStager, exchanged_batches

_EXPORT_KEYS = (
    "device", "human_res", "obj_res", "normal_res", "spatial_res", "spatial_grid", "spatial_indexgrid",
    "spatial_grid_metadata", "N_x", "N_y", "N_z", "spatial_occupancy_grids", "cache_count", "used_count",
    "principle_vec", "sub_principle_vec", "rel_dist_method", "rel_dist_thres", "normal_gaussian_sigma", "eps",
    "debug_obj_vert", "debug_obj_normal",
)

_STAGING_BYTES = 64 << 20


def _require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: coma_b200 runs on CUDA (sm_100a) only — no CPU fallback")


def load_voxelgrid(gridsize=3.0, resolution=24, center=[0, 0, 0]):
    """utils/coma_occupancy.py:160-183, expression for expression: the middle term `voxel_size * indexgrid.astype(f32)`
    is an fp32 product that numpy then promotes to fp64 — the bit-exact hit counts depend on it."""
    length_x = length_y = length_z = gridsize
    N_x = N_y = N_z = resolution
    voxel_size = gridsize / resolution
    center = np.array(center)
    start_point = center - np.array([length_x / 2, length_y / 2, length_z / 2])
    indexgrid = get_3d_indexgrid_ijk(N_x, N_y, N_z)
    canon_grid = start_point.reshape(3, 1, 1, 1) + voxel_size * indexgrid.astype(np.float32) + voxel_size / 2
    grid_metadata = dict(length_x=length_x, length_y=length_y, length_z=length_z, N_x=N_x, N_y=N_y, N_z=N_z,
                         start_point=start_point, voxel_size=voxel_size)
    return canon_grid, indexgrid, grid_metadata


class ComA_Occupancy:
    selected_obj_idxs = [0]

    def __init__(self, scale_tolerance: float, human_res: int, obj_res: int, normal_res: int, spatial_res: int,
                 proximity_settings=dict(), principle_vec=[0, 0, 1], sub_principle_vec=[0, 1, 0],
                 rel_dist_method: str = "dist", normal_gaussian_sigma: float = 0.1, selected_obj_idx: int = None,
                 eps: float = 1e-8, device: str = "cuda", human_slice=None):
        self.device = device
        self.human_res, self.obj_res = human_res, obj_res
        self.normal_res, self.spatial_res = normal_res, spatial_res
        assert normal_res == 0, "In this version, normal res is 0."

        grid, self.spatial_indexgrid, self.spatial_grid_metadata = load_voxelgrid(gridsize=2.4, resolution=self.spatial_res, center=[0, 0, 0])
        self.N_x, self.N_y, self.N_z = (self.spatial_grid_metadata[k] for k in ("N_x", "N_y", "N_z"))
        # per-axis centres (the [3,S,S,S] grid is separable); fp64, exactly the reference's values
        centers = np.stack([grid[0, :, 0, 0], grid[1, 0, :, 0], grid[2, 0, 0, :]])
        self._centers = torch.from_numpy(np.ascontiguousarray(centers)).to(device)
        self.spatial_grid = torch.from_numpy(grid).to(device)

        # B200 extension: this rank may own only a slice [h0, h1) of the human vertices (H-sharded multi-GPU)
        self._human_slice = (0, human_res) if human_slice is None else (int(human_slice[0]), int(human_slice[1]))
        h_local = self._human_slice[1] - self._human_slice[0]
        self.spatial_occupancy_grids = torch.zeros([h_local, self.N_x, self.N_y, self.N_z], dtype=torch.float32, device=device)

        self.cache_count = 0
        self.used_count = 0
        self.cache = dict()
        self.used = dict()

        self.principle_vec = torch.tensor(principle_vec, dtype=torch.float32).to(device)
        self.sub_principle_vec = torch.tensor(sub_principle_vec, dtype=torch.float32).to(device)

        assert rel_dist_method in ["dist", "sdf"], f"rel_dist_method: '{rel_dist_method}' not allowed"
        self.rel_dist_method = rel_dist_method
        self.rel_dist_thres = self.spatial_grid_metadata["voxel_size"] * scale_tolerance
        self.normal_gaussian_sigma = normal_gaussian_sigma
        self.eps = eps
        self.debug_obj_vert = None
        self.debug_obj_normal = None

    def register_sample_to_cache(self, **kwargs):
        self.cache[f"{self.cache_count:05}"] = kwargs
        self.cache_count = len(self.cache.keys())

    def aggregate_all_samples(self, exchange=False, group=None):
        """exchange=True (H-sharded multi-GPU): the cache holds only the samples THIS rank loaded; the canonicalised vertices
        are all-gathered so every rank scatters all samples into its rows, and `used_count` becomes the global count."""
        keys = list(self.cache.keys())
        n_global = self._aggregate_samples([self.cache[k] for k in keys], exchange=exchange, group=group)
        first = self.used_count
        for k in keys:
            self.used[f"{len(self.used):05}"] = self.cache[k]
        self.used_count = first + n_global
        self.cache = {}
        self.cache_count = 0

    def aggregate_single_sample(self, **kwargs):
        self._aggregate_samples([kwargs])

    def _canonical_human_verts(self, sample, all_rows=False, out=None):
        """Host part of aggregate_single_sample_for_occupancy (:274-288): invariants + subtraction in the input dtype.
        `out` (an fp32 row of the pinned staging buffer): the fp64 difference is rounded to fp32 on the store — the same values as
        the reference's `(human_verts - obj_vert).astype(float32)` without the temporary."""
        human_verts, obj_verts, obj_normals = sample["human_verts"], sample["obj_verts"], sample["obj_normals"]
        res = None
        for obj_idx in self.selected_obj_idxs:
            obj_vert, obj_normal = obj_verts[obj_idx], obj_normals[obj_idx]
            if self.debug_obj_vert is None:
                self.debug_obj_vert = obj_vert
            else:   # bitwise-equal objects (the normal case) skip np.allclose's ~15 us
                assert obj_vert is self.debug_obj_vert or np.array_equal(self.debug_obj_vert, obj_vert) or np.allclose(self.debug_obj_vert, obj_vert)
            if self.debug_obj_normal is None:
                self.debug_obj_normal = obj_normal
            else:
                assert obj_normal is self.debug_obj_normal or np.array_equal(self.debug_obj_normal, obj_normal) or np.allclose(self.debug_obj_normal, obj_normal)
            assert human_verts.shape[0] == self.human_res
            h0, h1 = (0, self.human_res) if all_rows else self._human_slice
            if out is None:
                res = human_verts[h0:h1] - obj_vert[None]     # only the rows this rank owns (elementwise: same values as slicing after)
            else:
                res = np.subtract(human_verts[h0:h1], obj_vert[None], out=out, casting="same_kind")
        return res

    def _stage_chunk(self, chunk, view, all_rows):
        """Fast path of `_canonical_human_verts` over a chunk of samples. False -> nothing is assumed written, use the numpy path."""
        from .staging import rows_equal_f64, stage_rows_f64
        if len(self.selected_obj_idxs) != 1 or not chunk:
            return False
        oi = self.selected_obj_idxs[0]
        try:
            hv = [s["human_verts"] for s in chunk]
            ov = [s["obj_verts"] for s in chunk]
            on = [s["obj_normals"] for s in chunk]
        except (KeyError, TypeError):
            return False
        if not all(isinstance(a, np.ndarray) for a in hv + ov + on) or any(a.shape[0] != self.human_res for a in hv):
            return False
        if any(a.ndim != 2 or a.shape[1] != 3 or a.shape[0] <= oi for a in ov + on):
            return False
        if self.debug_obj_vert is None:
            self.debug_obj_vert = ov[0][oi]
        if self.debug_obj_normal is None:
            self.debug_obj_normal = on[0][oi]
        # the reference's invariant (:277-284) on the normals: equal values is the normal case, anything else goes to np.allclose
        onr = on if oi == 0 else [a[oi:] for a in on]
        if not rows_equal_f64(onr, np.asarray(self.debug_obj_normal, dtype=np.float64)):
            return False
        h0, h1 = (0, self.human_res) if all_rows else self._human_slice
        ovr = ov if oi == 0 else [a[oi:] for a in ov]
        mism = stage_rows_f64(hv, view[:len(chunk)], row0=h0, sub=ovr, equal_to=np.asarray(self.debug_obj_vert, dtype=np.float64))
        return mism is not None and mism < 0

    def _aggregate_samples(self, samples, exchange=False, group=None):
        exchange = exchange and cdist.is_distributed(group)
        if not samples and not exchange:
            return 0
        _require_cuda(self.spatial_occupancy_grids, "ComA_Occupancy.aggregate")
        assert len(self.selected_obj_idxs) == 1
        h0, h1 = self._human_slice
        rows = self.human_res if exchange else max(h1 - h0, 1)
        chunk = max(32, min(8192, (_STAGING_BYTES // (rows * 12)) // 32 * 32))
        # <= 1024 samples per chunk: the host fill of chunk i + 1 overlaps the H2D copy and the K4 launch of chunk i (two buffers in flight)
        chunk = min(chunk, 256) if exchange else min(chunk, 1024, (len(samples) + 31) // 32 * 32)
        stager = BatchStager(dict(hvc=self.human_res if exchange else h1 - h0), chunk, self.spatial_occupancy_grids.device)
        outer = self

        class _Getter:   # writes the samples' canonical vertices straight into their pinned fp32 rows
            @staticmethod
            def fill(dst, i):
                outer._canonical_human_verts(samples[i], all_rows=exchange, out=dst)

            @staticmethod
            def fill_many(view, s0, n):
                """One library call per chunk (fp64 subtraction, fp32 store, same-object check) when every sample is a plain float64
                array; otherwise — or when the object differs bitwise from the first sample's — the per-sample numpy path decides."""
                if not outer._stage_chunk(samples[s0:s0 + n], view, all_rows=exchange):
                    for j in range(n):
                        outer._canonical_human_verts(samples[s0 + j], all_rows=exchange, out=view[j])
        getters = dict(hvc=_Getter)
        total = 0
        if exchange:
            for n, b in exchanged_batches(stager, getters, len(samples), group):
                if n and h1 > h0:
                    self.aggregate_batch_for_occupancy(b["hvc"][:, h0:h1].contiguous())
                total += n
        else:
            for n, b in stager.batches(getters, len(samples)):
                if h1 > h0:
                    self.aggregate_batch_for_occupancy(b["hvc"])
                total += n
        self.last_h2d_bytes = stager.h2d_bytes
        return total

    def aggregate_batch_for_occupancy(self, human_verts_canon):
        """Device-resident batched form of :289-295: human_verts_canon [S,H_local,3] fp32 (already minus obj vertex 0)."""
        ops.occupancy_accumulate(human_verts_canon, self._centers, self.rel_dist_thres, self.spatial_occupancy_grids)

    def normalize_prob_grid_for_spatials(self):
        """:297-300 (in place; NaN rows where a vertex never hit, as in the reference)."""
        ops.occupancy_readout(self.spatial_occupancy_grids, None)

    def normalize_prob_grid_for_spatials_v2(self):
        self.spatial_occupancy_grids = self.spatial_occupancy_grids / self.used_count

    def return_aggregated_spatial_grids(self, human_indices=None, group=None):
        """:305-312 -> torch tensor [N,N,N] on the device. With an H-sharded instance the per-rank fields are combined
        by one MAX all-reduce (NaN-propagating, like torch.max over the full vertex axis)."""
        sel = None
        h0, h1 = self._human_slice
        if human_indices is not None:
            idx = np.asarray(list(human_indices), dtype=np.int64).reshape(-1)
            if idx.size == 0:   # the reference's `grids[[]].max(dim=0)` raises on an empty selection
                raise IndexError("max(): Expected reduction dim 0 to have non-zero size (empty human_indices)")
            idx = np.where(idx < 0, idx + self.human_res, idx)
            idx = idx[(idx >= h0) & (idx < h1)] - h0
            # an H-sharded rank may own none of the selected vertices: an EMPTY selection (not "no selection") -> zero field,
            # the identity of the MAX all-reduce below (normalised occupancies are >= 0)
            sel = torch.tensor(idx, dtype=torch.int64, device=self.spatial_occupancy_grids.device)
        if h1 > h0:
            field = ops.occupancy_readout(self.spatial_occupancy_grids, sel)
        else:
            field = torch.zeros((self.N_x, self.N_y, self.N_z), dtype=torch.float32, device=self.spatial_occupancy_grids.device)
        sharded = self._human_slice != (0, self.human_res)
        return cdist.all_reduce_max_nan(field, group) if sharded else field

    def export(self, save_pth=None, group=None):
        """utils/coma_occupancy.py:314-330. H-sharded instance: a COLLECTIVE call — rank 0 assembles the full
        [H, Sg, Sg, Sg] grid on the host from the ranks' row blocks and returns / writes it; other ranks return None."""
        to_export = {}
        sharded = self._human_slice != (0, self.human_res) and cdist.is_distributed(group)
        for k in _EXPORT_KEYS:
            v = getattr(self, k)
            if sharded and k == "spatial_occupancy_grids":
                v = cdist.gather_rows(v, self.human_res, group, dst=0)
                if v is None:
                    continue
            if isinstance(v, torch.Tensor):
                v = v.detach().clone()
            elif isinstance(v, np.ndarray):
                v = v.copy()
            elif isinstance(v, dict):
                v = {kk: (vv.copy() if isinstance(vv, np.ndarray) else vv) for kk, vv in v.items()}
            to_export[k] = v
        if sharded:
            import torch.distributed as dist
            if dist.get_rank(group) != 0:
                return None
        to_export = to_np_torch_recursive(to_export, use_torch=False, device="cpu")
        if save_pth is None:
            return to_export
        with open(save_pth, "wb") as handle:
            pickle.dump(to_export, handle, protocol=pickle.HIGHEST_PROTOCOL)

    def load(self, load_pth):
        with open(load_pth, "rb") as handle:
            loadables = pickle.load(handle)
        h0, h1 = self._human_slice
        if self._human_slice != (0, self.human_res) and "spatial_occupancy_grids" in loadables:
            loadables["spatial_occupancy_grids"] = np.ascontiguousarray(loadables["spatial_occupancy_grids"][h0:h1])
        loadables = to_np_torch_recursive(loadables, use_torch=True, device=self.device)
        for k, v in loadables.items():
            setattr(self, k, v)


_all_reduce_max_nan = cdist.all_reduce_max_nan   # round-1 name, kept for callers / tests
