"""ComA accumulator — drop-in mirror of the reference's `utils/coma.py` class API (snuvclab/coma @ c89e2d1) whose hot
methods run as hand-written sm_100a kernels (coma_b200/csrc) through the C ABI in include/coma_b200.h.

Same constructor arguments, method names, error behaviour (AssertionError / NotImplementedError) and pickle layout as
the reference (`ComA`, utils/coma.py:176-610; `get_aggregated_contact` :614-641).  Differences, all B200-motivated:
  * samples registered in the cache are aggregated in BATCHES: the fp32 rounding of utils/misc.py:47-54 happens while
    filling pinned staging buffers, copies run on a side stream, and one K2 + one K3 launch consume a whole batch with
    the accumulators held in registers (the reference re-reads and re-writes every grid once per sample);
  * there is no CPU path: aggregation / read-outs on a non-CUDA device raise;
  * multi-GPU (one process per GPU): `ComA(..., human_slice=(h0, h1))` makes this rank own the rows [h0, h1) of every
    accumulator ("zero-collective" alternative). Every rank sees ALL samples — `aggregate_all_samples(exchange=True)`
    all-gathers the fp32 sample chunks each rank staged (a few hundred MB per job) — so the 31 GB histogram all-reduce of the
    sample-sharded form disappears; the only collectives left are the per-vertex ones of the read-outs ([O] flags, [O] / [H]
    maps). Sample-sharding + `all_reduce()` is kept for API parity with round 1.
"""
import pickle
from functools import partial

import numpy as np
import torch

from . import dist as cdist
from . import ops
from .misc import get_uniform_points_on_sphere, to_np_torch_recursive
from .staging import BatchStager, exchanged_batches


def negative_exp(x, spatial_grid_size, spatial_grid_thres, **kwargs):
    """Proximity score, utils/coma.py:116-119 (the threshold argument is unused there too)."""
    return torch.exp(-x / spatial_grid_size)


# The reference pickles `partial(utils.coma.negative_exp, ...)` BY REFERENCE inside every exported ComA
# (utils/coma.py:225,:584). Keep that module path so pickles travel in both directions (see utils/coma.py shim).
negative_exp.__module__ = "utils.coma"

_EXPORT_KEYS = (
    "device", "human_res", "obj_res", "normal_res", "spatial_res", "canon_normal_grid",
    "prob_grid_canon_human_wrt_obj", "prob_grid_canon_obj_wrt_human", "contact_dist_expectation_grid_nom",
    "contact_dist_expectation_grid_denom", "significant_contact_count", "proximity_settings", "contact_dist_func",
    "cross_contact_scores_nom", "cross_contact_scores_denom", "cache_count", "used_count", "principle_vec",
    "sub_principle_vec", "rel_dist_method", "normal_gaussian_sigma", "eps",
)

_STAGING_BYTES = 96 << 20  # pinned bytes per staging buffer set


def _require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: coma_b200 runs on CUDA (sm_100a) only — there is no CPU fallback "
                           f"(tensor is on '{t.device}')")


class _RowGetter:
    """Stager getter for one of the four per-sample arrays: `fill_many` converts a whole chunk (fp64 -> fp32, the rounding of
    utils/misc.py:47-54) with one multi-threaded call of the library's host helper when every array is a plain float64 [*, 3] array,
    and falls back to the per-sample numpy copy otherwise (other dtypes, non-contiguous views)."""

    def __init__(self, samples, key, rows):
        self.samples, self.key, self.rows = samples, key, rows

    def __call__(self, i):
        return self.samples[i][self.key][self.rows]

    def fill_many(self, view, s0, n):
        from .staging import stage_rows_f64
        arrs = [self.samples[s0 + j][self.key] for j in range(n)]
        row0 = self.rows.start or 0
        ok = all(isinstance(a, np.ndarray) for a in arrs) and stage_rows_f64(arrs, view[:n], row0=row0) is not None
        if not ok:
            for j in range(n):
                np.copyto(view[j], np.asarray(arrs[j])[self.rows], casting="same_kind")


class ComA:
    # Which device's arithmetic of the reference the bit-exact contact count follows: torch's CUDA reduction adds the squared
    # distance as (x2+z2)+y2, its CPU reduction as (x2+y2)+z2 (include/coma_b200.h). The reference's production runs are
    # device="cuda" (src/coma/extract_coma.py:329), hence the default; "cpu" reproduces a CPU run of the reference.
    reference_sum_order = "cuda"
    # K3: scores below 2^-orient_drop_bits (bins far outside the Gaussian's cone) are not evaluated; each bin stays within
    # S * 2^-32 absolute of the dense evaluation (include/coma_b200.h). 0 = dense kernel.
    orient_drop_bits = 32

    def __init__(self, human_res: int, obj_res: int, normal_res: int, spatial_res: int, proximity_settings=dict(),
                 principle_vec=[0, 0, 1], sub_principle_vec=[0, 1, 0], rel_dist_method: str = "dist",
                 normal_gaussian_sigma: float = 0.1, eps: float = 1e-8, device: str = "cuda", human_slice=None):
        self.device = device
        # B200 extension: this rank owns rows [h0, h1) of every accumulator (row-sharded multi-GPU); default = all rows
        self._human_slice = (0, human_res) if human_slice is None else (int(human_slice[0]), int(human_slice[1]))
        assert 0 <= self._human_slice[0] <= self._human_slice[1] <= human_res
        self.human_res, self.obj_res = human_res, obj_res
        self.normal_res, self.spatial_res = normal_res, spatial_res

        x, y, z = get_uniform_points_on_sphere(num_points=normal_res)
        grid_host = np.stack([x, y, z], axis=-1)
        self.canon_normal_grid = torch.tensor(grid_host).to(device)  # fp64 [N,3], :204-205
        # compact 32-bin patches of the bin centres for the cone-limited K3 kernel (host-side, once per instance)
        self._bin_perm = ops.bin_patches(grid_host, device) if 0 < normal_res <= 256 else None

        if self.spatial_res == 0:
            H, O, N = self._human_slice[1] - self._human_slice[0], obj_res, normal_res
            z32 = lambda *s: torch.zeros(s, dtype=torch.float32, device=device)
            self.prob_grid_canon_human_wrt_obj = z32(H, O, N)
            self.prob_grid_canon_obj_wrt_human = z32(H, O, N)
            self.contact_dist_expectation_grid_nom = z32(H, O)
            self.contact_dist_expectation_grid_denom = z32(H, O)
            self.significant_contact_count = z32(H, O)
        else:
            print("Please implement the spatial grid")
            raise NotImplementedError  # utils/coma.py:219-221

        self.proximity_settings = proximity_settings
        self.contact_dist_func = partial(negative_exp, **proximity_settings)
        self.cross_contact_scores_nom = torch.zeros([H, O], dtype=torch.float32, device=device)    # never written (:226)
        self.cross_contact_scores_denom = torch.zeros([H, O], dtype=torch.float32, device=device)

        self.cache_count = 0
        self.used_count = 0
        self.cache = dict()
        self.used = dict()

        self.principle_vec = torch.tensor(principle_vec, dtype=torch.float32).to(device)
        self.sub_principle_vec = torch.tensor(sub_principle_vec, dtype=torch.float32).to(device)

        assert rel_dist_method in ["dist", "sdf"], f"rel_dist_method: '{rel_dist_method}' not allowed"
        self.rel_dist_method = rel_dist_method
        self.normal_gaussian_sigma = normal_gaussian_sigma
        self.eps = eps

    # ------------------------------------------------------------------------------------------------- ingest
    def register_sample_to_cache(self, **kwargs):
        """Arrays are borrowed (held by reference, never copied or mutated), utils/coma.py:253-255."""
        self.cache[f"{self.cache_count:05}"] = kwargs
        self.cache_count = len(self.cache.keys())

    def aggregate_all_samples(self, exchange=False, group=None):
        """utils/coma.py:257-268 — but the whole cache goes through the GPU in a few batched launches.
        exchange=True (row-sharded multi-GPU): the cache holds only the samples THIS rank loaded; they are all-gathered so
        that every rank aggregates all samples into its rows, and `used_count` becomes the global sample count."""
        keys = list(self.cache.keys())
        samples = [self.cache[k] for k in keys]
        n_global = self._aggregate_samples(samples, exchange=exchange, group=group)
        first = self.used_count
        for k in keys:  # bookkeeping identical to the reference: cache -> used, counters
            self.used[f"{len(self.used):05}"] = self.cache[k]
        self.used_count = first + n_global
        self.cache = {}
        self.cache_count = 0

    def aggregate_single_sample(self, **kwargs):
        if self.spatial_res == 0:
            self._aggregate_samples([kwargs])   # NB like the reference, this does not touch `used_count` (:270-277)
        else:
            print("Please implement the spatial grid and aggregation in spatial grid")
            raise NotImplementedError

    def _aggregate_samples(self, samples, exchange=False, group=None):
        """-> number of samples aggregated (the global count in exchange mode)."""
        if self.rel_dist_method == "sdf":
            raise NotImplementedError  # utils/coma.py:325-326
        exchange = exchange and cdist.is_distributed(group)
        if not samples and not exchange:
            return 0
        for s in samples:
            self.assert_inputs(**s)
        _require_cuda(self.significant_contact_count, "ComA.aggregate")
        O = self.obj_res
        h0, h1 = self._human_slice
        Hs = self.human_res if exchange else h1 - h0        # rows staged per sample: all of them if they are to be exchanged
        rows = slice(None) if exchange else slice(h0, h1)
        per_sample = (Hs + O) * 3 * 4 * 2
        chunk = max(32, min(4096, (_STAGING_BYTES // per_sample) // 32 * 32))
        if exchange:
            chunk = min(chunk, 128)                          # world x chunk samples per launch
        else:
            # <= 128 samples per launch: the host fills the next pinned chunk (fp64 -> fp32, ~0.25 ms per sample at cfg 4) while
            # K2 / K3 of the previous one run; a K3 launch re-reads the grids, which costs ~3 % at 128 samples per launch
            chunk = min(chunk, 128, (len(samples) + 31) // 32 * 32)
        stager = BatchStager(dict(hv=Hs, hn=Hs, ov=O, on=O), chunk, self.significant_contact_count.device)
        getters = dict(hv=_RowGetter(samples, "human_verts", rows), hn=_RowGetter(samples, "human_normals", rows),
                       ov=_RowGetter(samples, "obj_verts", slice(None)), on=_RowGetter(samples, "obj_normals", slice(None)))
        total = 0
        if exchange:
            for n, b in exchanged_batches(stager, getters, len(samples), group):
                if n:
                    self.aggregate_batch_for_contact(b["hv"][:, h0:h1].contiguous(), b["hn"][:, h0:h1].contiguous(), b["ov"], b["on"])
                total += n
        else:
            for n, b in stager.batches(getters, len(samples)):
                self.aggregate_batch_for_contact(b["hv"], b["hn"], b["ov"], b["on"])
                total += n
        self.last_h2d_bytes = stager.h2d_bytes
        return total

    def aggregate_batch_for_contact(self, human_verts, human_normals, obj_verts, obj_normals):
        """Device-resident batched form of aggregate_single_sample_for_contact (utils/coma.py:279-323):
        human_* [S,H,3], obj_* [S,O,3] fp32 CUDA tensors. One K2 launch + one K3 launch for all S samples."""
        S = human_verts.shape[0]
        ps = self.proximity_settings
        ops.pair_accumulate(human_verts, obj_verts, ps["spatial_grid_thres"], ps["spatial_grid_size"],
                            self.significant_contact_count, self.contact_dist_expectation_grid_nom, sum_order=self.reference_sum_order)
        self.contact_dist_expectation_grid_denom += float(S)  # `+= 1.0` per sample (:291): a scalar in disguise
        grid = self.canon_normal_grid
        if grid.dtype != torch.float64:  # after load() the reference's grid is fp32 (:606)
            grid = grid.double()
        ops.orient_accumulate(human_normals, obj_normals, grid.contiguous(), self.normal_gaussian_sigma, self.eps,
                              self.principle_vec.tolist(), self.sub_principle_vec.tolist(),
                              self.prob_grid_canon_human_wrt_obj, self.prob_grid_canon_obj_wrt_human,
                              bin_perm=self._bin_perm, drop_bits=self.orient_drop_bits, sum_order=self.reference_sum_order)

    # ------------------------------------------------------------------------------------------------- read-outs
    def _contact_weights(self):
        dots = torch.sum(self.principle_vec[None, :] * self.canon_normal_grid, dim=-1)  # :342
        return ((1.0 - dots) / 2.0).to(torch.float32).contiguous()

    def normalize_prob_grid_for_normals(self):
        """utils/coma.py:328-330 (in place, re-applied on every call exactly like the reference)."""
        ops.normalize_contact_readout(self.prob_grid_canon_human_wrt_obj, self.eps)
        ops.normalize_contact_readout(self.prob_grid_canon_obj_wrt_human, self.eps)

    def compute_contact_map(self, contact_map_type: str, as_numpy: bool = True):
        """utils/coma.py:333-366. The normalisation pass and the weighted sum over bins are one fused kernel."""
        self.assert_inputs(contact_map_type=contact_map_type)
        _require_cuda(self.prob_grid_canon_human_wrt_obj, "ComA.compute_contact_map")
        w = self._contact_weights()
        nom, den = self.contact_dist_expectation_grid_nom, self.contact_dist_expectation_grid_denom
        want_h, want_o = contact_map_type in ["human", "both"], contact_map_type in ["obj", "both"]
        on_human = ops.normalize_contact_readout(self.prob_grid_canon_human_wrt_obj, self.eps, *((w, nom, den) if want_h else ()))
        on_obj = ops.normalize_contact_readout(self.prob_grid_canon_obj_wrt_human, self.eps, *((w, nom, den) if want_o else ()))
        contact_map_dict = {"human": on_human, "obj": on_obj}   # row-sharded instance: this rank's rows [h0, h1)
        if as_numpy:
            contact_map_dict = {k: (None if v is None else self._full_rows(v)) for k, v in contact_map_dict.items()}
            return to_np_torch_recursive(contact_map_dict, use_torch=False, device="cpu")
        return contact_map_dict

    # -- row-sharded instances: local blocks <-> full tensors (no-ops on an unsharded instance) --------------------------------
    def _sharded(self):
        return self._human_slice != (0, self.human_res)

    def _full_rows(self, t, group=None):
        return cdist.gather_rows(t, self.human_res, group) if self._sharded() else t

    def _significant(self, significant_contact_ratio, group=None):
        """-> (sig [H_local,O], any_o [H_local], any_h [O]); `any_h` is OR-ed over all ranks' rows on a sharded instance."""
        num = significant_contact_ratio * self.used_count
        sig, any_o, any_h = ops.significant_pairs(self.significant_contact_count, num)
        if self._sharded() and cdist.is_distributed(group):
            import torch.distributed as dist
            flags = any_h.to(torch.int32)
            dist.all_reduce(flags, op=dist.ReduceOp.MAX, group=group)
            any_h = flags > 0
        return sig, any_o, any_h

    def significant_contact_pairs(self, significant_contact_ratio: float, as_numpy: bool = True):
        """utils/coma.py:369-382."""
        sig = self._significant(significant_contact_ratio)[0]
        if as_numpy:
            return to_np_torch_recursive(self._full_rows(sig.to(torch.uint8)).to(torch.bool), use_torch=False, device="cpu")
        return sig

    def aggregate_contact_for_significant_pairs(self, contact_map_dict: dict, contact_map_type: str,
                                                significant_contact_ratio: float, as_numpy: bool = True):
        """utils/coma.py:385-438."""
        self.assert_inputs(contact_map_type=contact_map_type)
        sig, any_o, any_h = self._significant(significant_contact_ratio)  # any_o: [H] (OR over o); any_h: [O]
        agg_h = agg_o = None
        if contact_map_type in ["human", "both"]:
            assert contact_map_dict["human"] is not None, "If 'contact_map_type' is 'human' or 'both', contact_map_dict['human'] must not be None"
            agg_h = ops.masked_max(contact_map_dict["human"].to(torch.float32), any_h, axis=1)   # :407-413
        if contact_map_type in ["obj", "both"]:
            assert contact_map_dict["obj"] is not None, "If 'contact_map_type' is 'obj' or 'both', contact_map_dict['obj'] must not be None"
            agg_o = ops.masked_max(contact_map_dict["obj"].to(torch.float32), any_o, axis=0)     # :421-427
        if self._sharded():
            # the per-vertex maps are the only data that crosses GPUs: [H] rows gathered, [O] MAX-reduced (contact >= 0, and a
            # rank without significant rows contributes the reference's zero fallback, the identity of that max)
            agg_h = None if agg_h is None else self._full_rows(agg_h)
            agg_o = None if agg_o is None else cdist.all_reduce_max_nan(agg_o)
        out = {"human": agg_h, "obj": agg_o, "significant_contact_pairs": sig}
        if as_numpy:
            out["significant_contact_pairs"] = self._full_rows(sig.to(torch.uint8)).to(torch.bool)
            return to_np_torch_recursive(out, use_torch=False, device="cpu")
        return out

    def compute_nonphysical_response_sphere(self, n_bin: int, nonphysical_type: str, as_numpy: bool = True):
        """utils/coma.py:441-487."""
        self.assert_inputs(nonphysical_type=nonphysical_type)
        self.normalize_prob_grid_for_normals()
        sh = so = None
        if nonphysical_type in ["human", "both"]:
            sh = ops.entropy_readout(self.prob_grid_canon_human_wrt_obj, n_bin)
        if nonphysical_type in ["obj", "both"]:
            so = ops.entropy_readout(self.prob_grid_canon_obj_wrt_human, n_bin)
        out = {"human": sh, "obj": so, "n_bin": n_bin}
        if as_numpy:
            out = {k: (self._full_rows(v) if isinstance(v, torch.Tensor) else v) for k, v in out.items()}
            return to_np_torch_recursive(out, use_torch=False, device="cpu")
        return out

    def compute_nonphysical_response_sphere_v2(self, n_bin: int, nonphysical_type: str, as_numpy: bool = True):
        """utils/coma.py:529-579 (defined by the reference, used by none of its scripts): the per-bin normalised self-information
        weighted by the bin's alignment with the principle vector."""
        self.assert_inputs(nonphysical_type=nonphysical_type)
        self.normalize_prob_grid_for_normals()
        align = (self.canon_normal_grid * self.principle_vec[None]).sum(dim=-1).to(torch.float32).contiguous()   # :541
        sh = so = None
        if nonphysical_type in ["human", "both"]:
            sh = ops.entropy_readout(self.prob_grid_canon_human_wrt_obj, n_bin, weights=align)
        if nonphysical_type in ["obj", "both"]:
            so = ops.entropy_readout(self.prob_grid_canon_obj_wrt_human, n_bin, weights=align)
        out = {"human": sh, "obj": so, "n_bin": n_bin}
        if as_numpy:
            out = {k: (self._full_rows(v) if isinstance(v, torch.Tensor) else v) for k, v in out.items()}
            return to_np_torch_recursive(out, use_torch=False, device="cpu")
        return out

    def assert_inputs(self, **kwargs):
        """utils/coma.py:489-526."""
        for key, res in (("human_verts", self.human_res), ("human_normals", self.human_res),
                         ("obj_verts", self.obj_res), ("obj_normals", self.obj_res)):
            if key in kwargs:
                v = kwargs[key]
                assert v.ndim == 2
                assert v.shape[-1] == 3
                assert len(v) == res
        if "contact_map_type" in kwargs:
            assert kwargs["contact_map_type"] in ["human", "obj", "both"], "Only ['human'/'obj'/'both'] allowed for Argument: 'contact_map_type'"
        if "nonphysical_type" in kwargs:
            assert kwargs["nonphysical_type"] in ["human", "obj", "both"], "Only ['human'/'obj'/'both'] allowed for Argument: 'nonphysical_type'"

    # ------------------------------------------------------------------------------------------------- multi-GPU
    def all_reduce(self, group=None):
        """Sum the accumulators of sample-sharded ranks: one all-reduce over NVLink per tensor.
        Integer-valued counts stay exact (< 2^24); `denom` / `used_count` become the global number of samples."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        assert not self._sharded(), "all_reduce() sums SAMPLE-sharded accumulators; a row-sharded instance needs no reduction"
        for name in ("significant_contact_count", "contact_dist_expectation_grid_nom",
                     "contact_dist_expectation_grid_denom", "prob_grid_canon_human_wrt_obj",
                     "prob_grid_canon_obj_wrt_human"):
            dist.all_reduce(getattr(self, name), op=dist.ReduceOp.SUM, group=group)
        dev = self.significant_contact_count.device
        n = torch.tensor([self.used_count], dtype=torch.int64, device=dev)
        dist.all_reduce(n, op=dist.ReduceOp.SUM, group=group)
        self.used_count = int(n.item())

    # ------------------------------------------------------------------------------------------------- persistence
    _ROW_SHARDED_KEYS = ("prob_grid_canon_human_wrt_obj", "prob_grid_canon_obj_wrt_human", "contact_dist_expectation_grid_nom",
                         "contact_dist_expectation_grid_denom", "significant_contact_count", "cross_contact_scores_nom",
                         "cross_contact_scores_denom")

    def export(self, save_pth=None, group=None):
        """utils/coma.py:582-597: every tensor -> numpy fp32/int64, same keys, grids stored un-normalised.
        Row-sharded instance: a COLLECTIVE call — every rank sends its row blocks to rank 0, which assembles the full
        [H, ...] arrays on the host and returns / writes them; the other ranks return None."""
        to_export = {}
        sharded = self._sharded() and cdist.is_distributed(group)
        for k in _EXPORT_KEYS:
            v = getattr(self, k)
            if sharded and k in self._ROW_SHARDED_KEYS:
                v = cdist.gather_rows(v, self.human_res, group, dst=0)
            to_export[k] = v.detach().clone() if isinstance(v, torch.Tensor) else v
        if sharded:
            import torch.distributed as dist
            if dist.get_rank(group) != 0:
                return None
        to_export["proximity_settings"] = dict(self.proximity_settings)
        to_export["contact_dist_func"] = _portable_partial(self.proximity_settings)
        to_export = to_np_torch_recursive(to_export, use_torch=False, device="cpu")
        if save_pth is None:
            return to_export
        with open(save_pth, "wb") as handle:
            pickle.dump(to_export, handle, protocol=pickle.HIGHEST_PROTOCOL)

    def load(self, load_pth):
        """utils/coma.py:600-610."""
        with open(load_pth, "rb") as handle:
            loadables = pickle.load(handle)
        h0, h1 = self._human_slice
        if self._sharded():   # keep only this rank's rows (sliced on the host, before anything reaches the device)
            for k in self._ROW_SHARDED_KEYS:
                if k in loadables:
                    loadables[k] = np.ascontiguousarray(loadables[k][h0:h1])
        loadables = to_np_torch_recursive(loadables, use_torch=True, device=self.device)
        for k, v in loadables.items():
            setattr(self, k, v)


def _portable_partial(proximity_settings):
    """partial(utils.coma.negative_exp, ...) bound to whatever `utils.coma` resolves to in this process, so the pickle
    is loadable by the reference and by this package alike."""
    try:
        import utils.coma as uc
        fn = uc.negative_exp
    except Exception:  # no `utils` package on sys.path: pickle a reference to this module instead
        fn = negative_exp
    return partial(fn, **proximity_settings)


def get_aggregated_contact(coma: ComA, contact_map_type: str, significant_contact_ratio: float):
    """utils/coma.py:614-641 -> (np.float32 [H or O], np.int64 [k])."""
    assert contact_map_type in ["human", "obj"]
    contact_map_dict = coma.compute_contact_map(contact_map_type=contact_map_type, as_numpy=False)
    aggregated_contact_dict = coma.aggregate_contact_for_significant_pairs(
        contact_map_dict=contact_map_dict, contact_map_type=contact_map_type,
        significant_contact_ratio=significant_contact_ratio, as_numpy=True)
    aggregated_contact = aggregated_contact_dict[contact_map_type]
    significant_contact_pairs = aggregated_contact_dict["significant_contact_pairs"]
    indicator = np.any(significant_contact_pairs, axis=0 if contact_map_type == "human" else 1)
    significant_contact_vertex_indices = np.argwhere(indicator)[:, 0]
    return aggregated_contact, significant_contact_vertex_indices


def get_nonphysical_score(coma: ComA, nonphysical_type: str):
    """utils/coma.py:645-646."""
    return coma.compute_nonphysical_response_sphere(n_bin=1e6, nonphysical_type=nonphysical_type, as_numpy=True)[nonphysical_type]


def simplify_mesh_and_get_indices(mesh, number_of_points: int, simplify_method="poisson_disk", mesh_index_find_method="distance-based",
                                  debug=False, device="cuda"):
    """utils/coma.py:29-98. `mesh` is whatever the caller samples from — an open3d TriangleMesh in the reference's scripts
    (src/coma/downsample_human.py, downsample_objects.py), or any object with `.vertices` and the two sampling methods
    `sample_points_poisson_disk(number_of_points=)` / `sample_points_uniformly(number_of_points=)` returning an object with `.points`
    (open3d itself is not a dependency of this package: the SAMPLING is the caller's, the O(V*N) index search is K1 on the GPU,
    fp64-exact with np.argmin's first-minimum rule). Only the distance-based index finder works in the reference (its ray-tracing
    branch drops into an IPython shell, :60-68) and only that one exists here. Returns (list of int vertex indices, the sampled pcd)."""
    if simplify_method == "poisson_disk":
        pcd = mesh.sample_points_poisson_disk(number_of_points=number_of_points)
    elif simplify_method == "uniform":
        pcd = mesh.sample_points_uniformly(number_of_points=number_of_points)
    else:
        raise NotImplementedError
    if mesh_index_find_method != "distance-based":
        raise NotImplementedError
    return nearest_vertex_indices(np.asarray(pcd.points), np.asarray(mesh.vertices), device=device), pcd


def nearest_vertex_indices(points, mesh_verts, device="cuda"):
    """The distance-based index finder of simplify_mesh_and_get_indices (utils/coma.py:87-91, :96):
    points [N,3], mesh_verts [V,3] (any float dtype, host) -> list of int vertex indices (fp64-exact argmin)."""
    pts = torch.tensor(np.asarray(points, dtype=np.float64), device=device)
    verts = torch.tensor(np.asarray(mesh_verts, dtype=np.float64), device=device)
    return list(ops.nearest_vertex(pts, verts).cpu().numpy())
