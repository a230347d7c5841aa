"""TEST INFRASTRUCTURE: oracle-backed CPU stand-ins for the CUDA entry points of `coma_b200.ops` / `coma_b200.ingest`.

The product has no CPU path (the classes raise on a non-CUDA device). The multi-process tests that run in the GPU-less
container (gloo, world_size 2) exercise the HOST logic around the kernels — row sharding, the sample exchange, the
collective read-outs, export / load of sharded instances, the CLI's rank choreography — so they swap the kernel calls for the
oracle's restatements of the same functions. Nothing outside tests/ imports this module.
"""
import numpy as np
import torch


def install():
    """Patch the kernel entry points; returns a callable that restores the originals (the product must stay CUDA-only for
    every other test in the process)."""
    from coma_b200 import coma as coma_mod
    from coma_b200 import coma_occupancy as occ_mod
    from coma_b200 import ingest, ops
    from oracle import oracle

    def _np(t):
        return t.detach().cpu().numpy()

    def pair_accumulate(hv, ov, thres, grid_size, count, nom, sum_order="cpu"):
        c, n = oracle.pair_accumulate(_np(hv), _np(ov), thres, grid_size, sum_order=sum_order)
        count += torch.from_numpy(c)
        nom += torch.from_numpy(n)

    def orient_accumulate(hn, on, grid, sigma, eps, p, sub_p, PH, PO, bin_perm=None, drop_bits=0, sum_order="cpu"):
        a, b = oracle.orient_accumulate(_np(hn), _np(on), _np(grid), sigma, eps, p, sub_p, sum_order=sum_order)
        PH += torch.from_numpy(a)
        PO += torch.from_numpy(b)

    def normalize_contact_readout(P, eps, w=None, nom=None, denom=None):
        Pn = torch.from_numpy(oracle.normalize_normals(_np(P), eps))
        P.copy_(Pn)
        if w is None:
            return None
        return ((Pn * w[None, None, :]).sum(-1) * (nom / denom)).float()

    def significant_pairs(count, num):
        sig = count >= num
        return sig, sig.any(1), sig.any(0)

    def masked_max(cmap, mask, axis):
        mask = mask.bool()
        if axis == 1:
            return cmap[:, mask].max(-1).values if mask.any() else torch.zeros(cmap.shape[0])
        return cmap[mask, :].max(0).values if mask.any() else torch.zeros(cmap.shape[1])

    def entropy_readout(P, n_bin):
        return torch.from_numpy(oracle.entropy_score(_np(P), n_bin))

    def occupancy_accumulate(hvc, centers, thr, grids):
        S, H, _ = hvc.shape
        Sg = centers.shape[1]
        g = np.zeros((H, Sg, Sg, Sg), np.float32)
        hv = _np(hvc).astype(np.float64)
        oracle.lib().oracle_occupancy_accumulate(oracle._p(np.ascontiguousarray(hv, np.float32), oracle._f32p), S, H,
                                                 oracle._p(np.ascontiguousarray(_np(centers)), oracle._f64p), Sg, float(thr),
                                                 oracle._p(g, oracle._f32p))
        grids += torch.from_numpy(g)

    def occupancy_readout(grids, sel_idx=None):
        field, norm = oracle.occupancy_field(_np(grids))
        grids.copy_(torch.from_numpy(norm))
        if sel_idx is not None:
            idx = _np(sel_idx)
            if idx.size == 0:
                return torch.zeros(grids.shape[1:])
            sub = norm[idx]
            field = np.where(np.isnan(sub).any(0), np.float32(np.nan), sub.max(0)).astype(np.float32)
        return torch.from_numpy(field)

    saved = []

    def patch(obj, name, value):
        saved.append((obj, name, getattr(obj, name)))
        setattr(obj, name, value)

    for name, fn in dict(pair_accumulate=pair_accumulate, orient_accumulate=orient_accumulate,
                         normalize_contact_readout=normalize_contact_readout, significant_pairs=significant_pairs,
                         masked_max=masked_max, entropy_readout=entropy_readout, occupancy_accumulate=occupancy_accumulate,
                         occupancy_readout=occupancy_readout).items():
        patch(ops, name, fn)
    patch(coma_mod, "_require_cuda", lambda t, what: None)

    class CpuMeshNormals:
        def __init__(self, faces, num_verts, device="cpu"):
            self.faces, self.V, self.dev = np.asarray(faces), num_verts, torch.device("cpu")

        def __call__(self, verts, eps=-1.0):
            v = np.asarray(verts, dtype=np.float64)
            out = oracle.vertex_normals(v.reshape(-1, self.V, 3), self.faces, None if eps < 0 else eps)
            return torch.from_numpy(out.reshape(v.shape))

    patch(ingest, "mesh_normals_for", lambda faces, num_verts, device="cpu": CpuMeshNormals(faces, num_verts))
    patch(ingest, "vertex_normals",
          lambda verts, faces, eps=-1.0, device="cpu": CpuMeshNormals(faces, np.asarray(verts).shape[-2])(verts, eps).numpy())
    patch(occ_mod, "_require_cuda", lambda t, what: None)

    def restore():
        for obj, name, value in reversed(saved):
            setattr(obj, name, value)
    return restore
