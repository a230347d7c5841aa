"""Per-kernel timings on one GPU (CUDA events, warm-up, L2 flush between iterations). Development tool; the judged
numbers come from bench.py."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coma_b200 import ops, synth  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main():
    H, O, N = 10475, 1500, 250
    out = {}
    for S in (1, 8, 64):
        hv, hn, ov, on = (torch.from_numpy(a).to(dev) for a in synth.make_sample_arrays(S, H, O, seed=1))
        count, nom = torch.zeros((H, O), device=dev), torch.zeros((H, O), device=dev)
        ms = timeit(lambda: ops.pair_accumulate(hv, ov, 0.05, 0.15, count, nom))
        out[f"k2_S{S}"] = dict(ms=ms, pairs_per_s=S * H * O / ms * 1e3, gbs_16B=16 * H * O / ms / 1e6)
    grid = torch.tensor(np.stack(__import__("coma_b200.misc", fromlist=["x"]).get_uniform_points_on_sphere(N), -1), device=dev)
    PH, PO = torch.zeros((H, O, N), device=dev), torch.zeros((H, O, N), device=dev)
    for variant in ("v1", "x2"):
        os.environ["COMA_B200_K3"] = variant
        for S in (1, 64):
            hv, hn, ov, on = (torch.from_numpy(a).to(dev) for a in synth.make_sample_arrays(S, H, O, seed=1))
            ms = timeit(lambda: ops.orient_accumulate(hn, on, grid, 0.25, 1e-10, [0, 0, 1], [0, 1, 0], PH, PO), iters=3, warm=1)
            out[f"k3_{variant}_S{S}"] = dict(ms=ms, pair_samples_per_s=S * H * O / ms * 1e3, bin_evals_per_s=2 * N * S * H * O / ms * 1e3,
                                             gbs_rmw=16 * H * O * N / ms / 1e6)
    os.environ["COMA_B200_K3"] = "x2"
    del PH, PO
    for Sg, S in ((30, 256), (128, 64)):
        Hh = 10475 if Sg == 30 else 2048
        hv, hn, ov, on = synth.make_sample_arrays(S, Hh, 4, seed=2)
        hvc = torch.from_numpy((hv - ov[:, 0:1]).astype(np.float32)).to(dev)
        from coma_b200.coma_occupancy import load_voxelgrid
        g, _, meta = load_voxelgrid(2.4, Sg)
        centers = torch.from_numpy(np.ascontiguousarray(np.stack([g[0, :, 0, 0], g[1, 0, :, 0], g[2, 0, 0, :]]))).to(dev)
        grids = torch.zeros((Hh, Sg, Sg, Sg), device=dev)
        # (A/B against round 1's kernels: tools/occ_ab.py — the COMA_B200_OCC_PATH / COMA_B200_K5C switches are read once per process)
        grids.zero_()
        ms = timeit(lambda: ops.occupancy_accumulate(hvc, centers, meta["voxel_size"] * 3.0, grids), iters=3, warm=1)
        hits = grids.sum().item() / 4
        out[f"k4_Sg{Sg}"] = dict(ms=ms, vertex_samples_per_s=S * Hh / ms * 1e3, hits_per_vs=hits / (S * Hh))
        ms = timeit(lambda: ops.occupancy_readout(grids, None), iters=2, warm=1)
        out[f"k5c_Sg{Sg}"] = dict(ms=ms, dense_equivalent_gbs=3 * 4 * Hh * Sg**3 / ms / 1e6)   # (re-normalising already normalised grids)
        del grids
    Hs, Os = 4000, 1500
    P = torch.rand((Hs, Os, N), device=dev)
    w = torch.rand(N, device=dev)
    nom, den = torch.rand((Hs, Os), device=dev), torch.ones((Hs, Os), device=dev)
    ms = timeit(lambda: ops.normalize_contact_readout(P, 1e-10, w, nom, den), iters=3, warm=1)
    out["k5a"] = dict(ms=ms, gbs=8 * Hs * Os * N / ms / 1e6)
    ms = timeit(lambda: ops.entropy_readout(P, 1e6), iters=3, warm=1)
    out["k5b"] = dict(ms=ms, gbs=4 * Hs * Os * N / ms / 1e6)
    verts = torch.randn((10475, 3), dtype=torch.float64, device=dev)
    pts = torch.randn((2048, 3), dtype=torch.float64, device=dev)
    ms = timeit(lambda: ops.nearest_vertex(pts, verts))
    out["k1"] = dict(ms=ms, pairs_per_s=2048 * 10475 / ms * 1e3)
    from coma_b200.ingest import MeshNormals
    rng = np.random.default_rng(0)
    Vn, Fn, Sn = 10475, 20908, 64                       # SMPL-X sized topology (random connectivity), 64 samples per launch
    faces = rng.integers(0, Vn, (Fn, 3))
    mn = MeshNormals(faces, Vn, dev)
    vb = torch.randn((Sn, Vn, 3), dtype=torch.float64, device=dev)
    ms = timeit(lambda: mn(vb, 1e-10))
    out["k6"] = dict(ms=ms, meshes_per_s=Sn / ms * 1e3, vertices_per_s=Sn * Vn / ms * 1e3)
    print(json.dumps(out, indent=1))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/microbench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
